from __future__ import annotations

import dataclasses

from d9d_b200.loop.event import Event

from .common import (
    EventConfigurationStartedContext,
    EventDataLoaderReadyContext,
    EventModelStagesReadyContext,
    EventStepContext,
)


@dataclasses.dataclass(kw_only=True)
class EventInferenceReadyContext:
    pass


@dataclasses.dataclass(kw_only=True)
class EventInferenceFinishedContext:
    pass


EVENT_INFERENCE_CONFIG_STARTED = Event[EventConfigurationStartedContext](id="inference.configuration.start")
EVENT_INFERENCE_DATA_LOADER_READY = Event[EventDataLoaderReadyContext](id="inference.configuration.data_loader")
EVENT_INFERENCE_MODEL_STAGES_READY = Event[EventModelStagesReadyContext](id="inference.configuration.model_stages")
EVENT_INFERENCE_READY = Event[EventInferenceReadyContext](id="inference.ready")
EVENT_INFERENCE_STEP_PRE = Event[EventStepContext](id="inference.step.pre")
EVENT_INFERENCE_STEP_POST = Event[EventStepContext](id="inference.step.post")
EVENT_INFERENCE_FORWARD_PRE = Event[EventStepContext](id="inference.forward.pre")
EVENT_INFERENCE_FORWARD_POST = Event[EventStepContext](id="inference.forward.post")
EVENT_INFERENCE_FINISHED = Event[EventInferenceFinishedContext](id="inference.finished")
