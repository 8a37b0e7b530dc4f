"""Events of ``Trainer`` in firing order: five configuration events, then per step ``step.pre`` →
``forward_backward.pre/post`` → ``optimizer_step.pre/post`` → ``step.post``, finally ``finished``."""

from __future__ import annotations

from .common import namespaced
from .contexts import (
    EventConfigurationStartedContext,
    EventDataLoaderReadyContext,
    EventLRSchedulerReadyContext,
    EventModelStagesReadyContext,
    EventOptimizerReadyContext,
    EventStepContext,
    EventTrainFinishedContext,
    EventTrainReadyContext,
)

_train = namespaced("train")

EVENT_TRAIN_CONFIG_STARTED = _train("configuration.start", EventConfigurationStartedContext)
EVENT_TRAIN_DATA_LOADER_READY = _train("configuration.data_loader", EventDataLoaderReadyContext)
EVENT_TRAIN_MODEL_STAGES_READY = _train("configuration.model_stages", EventModelStagesReadyContext)
EVENT_TRAIN_OPTIMIZER_READY = _train("configuration.optimizer", EventOptimizerReadyContext)
EVENT_TRAIN_LR_SCHEDULER_READY = _train("configuration.lr_scheduler", EventLRSchedulerReadyContext)

EVENT_TRAIN_READY = _train("ready", EventTrainReadyContext)
EVENT_TRAIN_STEP_PRE, EVENT_TRAIN_STEP_POST = (_train(f"step.{edge}", EventStepContext) for edge in ("pre", "post"))
EVENT_TRAIN_FORWARD_BACKWARD_PRE, EVENT_TRAIN_FORWARD_BACKWARD_POST = (
    _train(f"forward_backward.{edge}", EventStepContext) for edge in ("pre", "post"))
EVENT_TRAIN_OPTIMIZER_STEP_PRE, EVENT_TRAIN_OPTIMIZER_STEP_POST = (
    _train(f"optimizer_step.{edge}", EventStepContext) for edge in ("pre", "post"))
EVENT_TRAIN_FINISHED = _train("finished", EventTrainFinishedContext)

__all__ = [name for name in dir() if name.startswith(("EVENT_TRAIN_", "Event"))]
