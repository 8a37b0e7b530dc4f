from __future__ import annotations

import abc
from typing import Any, Generic, TypeVar

import torch
from torch.distributed.checkpoint.stateful import Stateful

from d9d_b200.core.dist_context import DistributedContext

TComputeResult = TypeVar("TComputeResult")


class Metric(abc.ABC, Stateful, Generic[TComputeResult]):
    """Statistic accumulated over steps, synchronisable across ranks and checkpointable."""

    @abc.abstractmethod
    def update(self, *args: Any, **kwargs: Any) -> None: ...

    @abc.abstractmethod
    def sync(self, dist_context: DistributedContext) -> None: ...

    @abc.abstractmethod
    def compute(self) -> TComputeResult: ...

    @abc.abstractmethod
    def reset(self) -> None: ...

    def to(self, device: str | torch.device | int) -> None:  # noqa: B027
        """Move the state to ``device`` (no-op for stateless metrics)."""
