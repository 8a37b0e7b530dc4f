from __future__ import annotations

import dataclasses
from typing import Protocol, runtime_checkable

from torch import nn
from torch.optim import Optimizer

from d9d_b200.core.dist_context import DistributedContext


@dataclasses.dataclass(kw_only=True)
class InitializeOptimizerStageContext:
    dist_context: DistributedContext
    model: nn.Module


@runtime_checkable
class OptimizerProvider(Protocol):
    def __call__(self, context: InitializeOptimizerStageContext) -> Optimizer: ...
