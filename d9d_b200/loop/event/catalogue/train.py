from __future__ import annotations

import dataclasses

from d9d_b200.core.protocol import LRSchedulerProtocol, OptimizerProtocol
from d9d_b200.loop.event import Event
from d9d_b200.tracker import BaseTrackerRun

from .common import (
    EventConfigurationStartedContext,
    EventDataLoaderReadyContext,
    EventModelStagesReadyContext,
    EventStepContext,
)


@dataclasses.dataclass(kw_only=True)
class EventOptimizerReadyContext:
    optimizer: OptimizerProtocol


@dataclasses.dataclass(kw_only=True)
class EventLRSchedulerReadyContext:
    lr_scheduler: LRSchedulerProtocol


@dataclasses.dataclass(kw_only=True)
class EventTrainReadyContext:
    run: BaseTrackerRun


@dataclasses.dataclass(kw_only=True)
class EventTrainFinishedContext:
    pass


# configuration
EVENT_TRAIN_CONFIG_STARTED = Event[EventConfigurationStartedContext](id="train.configuration.start")
EVENT_TRAIN_DATA_LOADER_READY = Event[EventDataLoaderReadyContext](id="train.configuration.data_loader")
EVENT_TRAIN_MODEL_STAGES_READY = Event[EventModelStagesReadyContext](id="train.configuration.model_stages")
EVENT_TRAIN_OPTIMIZER_READY = Event[EventOptimizerReadyContext](id="train.configuration.optimizer")
EVENT_TRAIN_LR_SCHEDULER_READY = Event[EventLRSchedulerReadyContext](id="train.configuration.lr_scheduler")
# runtime
EVENT_TRAIN_READY = Event[EventTrainReadyContext](id="train.ready")
EVENT_TRAIN_STEP_PRE = Event[EventStepContext](id="train.step.pre")
EVENT_TRAIN_STEP_POST = Event[EventStepContext](id="train.step.post")
EVENT_TRAIN_FORWARD_BACKWARD_PRE = Event[EventStepContext](id="train.forward_backward.pre")
EVENT_TRAIN_FORWARD_BACKWARD_POST = Event[EventStepContext](id="train.forward_backward.post")
EVENT_TRAIN_OPTIMIZER_STEP_PRE = Event[EventStepContext](id="train.optimizer_step.pre")
EVENT_TRAIN_OPTIMIZER_STEP_POST = Event[EventStepContext](id="train.optimizer_step.post")
EVENT_TRAIN_FINISHED = Event[EventTrainFinishedContext](id="train.finished")
