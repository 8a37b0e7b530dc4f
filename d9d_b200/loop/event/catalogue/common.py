from __future__ import annotations

import dataclasses
from typing import TYPE_CHECKING

from torch import nn
from torch.utils.data import DataLoader

from d9d_b200.core.dist_context import DistributedContext

if TYPE_CHECKING:
    from d9d_b200.loop.component import Stepper


@dataclasses.dataclass(kw_only=True)
class EventStepContext:
    stepper: "Stepper"


@dataclasses.dataclass(kw_only=True)
class EventConfigurationStartedContext:
    dist_context: DistributedContext


@dataclasses.dataclass(kw_only=True)
class EventDataLoaderReadyContext:
    data_loader: DataLoader


@dataclasses.dataclass(kw_only=True)
class EventModelStagesReadyContext:
    modules: list[nn.Module]
