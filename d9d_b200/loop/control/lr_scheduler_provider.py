from __future__ import annotations

import dataclasses
from typing import Protocol, runtime_checkable

from torch.optim import Optimizer

from d9d_b200.core.dist_context import DistributedContext
from d9d_b200.core.protocol import LRSchedulerProtocol


@dataclasses.dataclass(kw_only=True)
class InitializeLRSchedulerContext:
    dist_context: DistributedContext
    total_steps: int
    optimizer: Optimizer


@runtime_checkable
class LRSchedulerProvider(Protocol):
    def __call__(self, context: InitializeLRSchedulerContext) -> LRSchedulerProtocol: ...
