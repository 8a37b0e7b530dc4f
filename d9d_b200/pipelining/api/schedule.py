from __future__ import annotations

import abc
from typing import Any

import torch

from .sharding import PipelineShardingSpec


class PipelineSchedule(abc.ABC):
    """Executes one training / inference step over all microbatches of a global batch."""

    @abc.abstractmethod
    def configure_buffers(self, inputs: dict[str, torch.Tensor], kwargs: dict[str, Any],
                          sharding_spec: PipelineShardingSpec | None) -> None:
        """Plan microbatch sharding and communication buffers for inputs of this structure."""

    @abc.abstractmethod
    def step(self, inputs: dict[str, torch.Tensor], kwargs: dict[str, Any]) -> None:
        """Run forward (and backward) for every microbatch according to the schedule."""
