"""Events of ``Inference``: three configuration events, ``ready``, per batch ``step.pre`` → ``forward.pre/post`` →
``step.post``, and ``finished``."""

from __future__ import annotations

from .common import namespaced
from .contexts import (
    EventConfigurationStartedContext,
    EventDataLoaderReadyContext,
    EventInferenceFinishedContext,
    EventInferenceReadyContext,
    EventModelStagesReadyContext,
    EventStepContext,
)

_infer = namespaced("inference")

EVENT_INFERENCE_CONFIG_STARTED = _infer("configuration.start", EventConfigurationStartedContext)
EVENT_INFERENCE_DATA_LOADER_READY = _infer("configuration.data_loader", EventDataLoaderReadyContext)
EVENT_INFERENCE_MODEL_STAGES_READY = _infer("configuration.model_stages", EventModelStagesReadyContext)
EVENT_INFERENCE_READY = _infer("ready", EventInferenceReadyContext)
EVENT_INFERENCE_STEP_PRE, EVENT_INFERENCE_STEP_POST = (_infer(f"step.{edge}", EventStepContext) for edge in ("pre", "post"))
EVENT_INFERENCE_FORWARD_PRE, EVENT_INFERENCE_FORWARD_POST = (_infer(f"forward.{edge}", EventStepContext) for edge in ("pre", "post"))
EVENT_INFERENCE_FINISHED = _infer("finished", EventInferenceFinishedContext)

__all__ = [name for name in dir() if name.startswith(("EVENT_INFERENCE_", "Event"))]
