"""Payload types of the loop events (one place for both the training and the inference catalogue)."""

from __future__ import annotations

import dataclasses
from typing import TYPE_CHECKING

from torch import nn
from torch.utils.data import DataLoader

from d9d_b200.core.dist_context import DistributedContext
from d9d_b200.core.protocol import LRSchedulerProtocol, OptimizerProtocol
from d9d_b200.tracker import BaseTrackerRun

if TYPE_CHECKING:
    from d9d_b200.loop.component import Stepper


def _payload(cls: type) -> type:
    return dataclasses.dataclass(kw_only=True)(cls)


@_payload
class EventConfigurationStartedContext:
    dist_context: DistributedContext


@_payload
class EventDataLoaderReadyContext:
    data_loader: DataLoader


@_payload
class EventModelStagesReadyContext:
    modules: list[nn.Module]


@_payload
class EventOptimizerReadyContext:
    optimizer: OptimizerProtocol


@_payload
class EventLRSchedulerReadyContext:
    lr_scheduler: LRSchedulerProtocol


@_payload
class EventStepContext:
    stepper: "Stepper"


@_payload
class EventTrainReadyContext:
    run: BaseTrackerRun


@_payload
class EventTrainFinishedContext:
    pass


@_payload
class EventInferenceReadyContext:
    pass


@_payload
class EventInferenceFinishedContext:
    pass
