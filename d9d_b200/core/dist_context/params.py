from __future__ import annotations

import logging
from typing import TYPE_CHECKING, Self

from pydantic import BaseModel, ConfigDict, model_validator

if TYPE_CHECKING:
    from .configured import DistributedContext


class DeviceMeshParameters(BaseModel):
    """Degrees of every parallelism axis. World size = product of all but ``expert_parallel``."""

    model_config = ConfigDict(frozen=True)

    pipeline_parallel: int = 1
    data_parallel_replicate: int = 1
    data_parallel_shard: int = 1
    context_parallel_replicate: int = 1
    context_parallel_shard: int = 1
    tensor_parallel: int = 1
    expert_parallel: int = 1

    @property
    def has_pipeline_parallel(self) -> bool:
        return self.pipeline_parallel > 1

    @property
    def has_data_parallel_replicate(self) -> bool:
        return self.data_parallel_replicate > 1

    @property
    def has_data_parallel_shard(self) -> bool:
        return self.data_parallel_shard > 1

    @property
    def has_context_parallel_replicate(self) -> bool:
        return self.context_parallel_replicate > 1

    @property
    def has_context_parallel_shard(self) -> bool:
        return self.context_parallel_shard > 1

    @property
    def has_tensor_parallel(self) -> bool:
        return self.tensor_parallel > 1

    @property
    def has_expert_parallel(self) -> bool:
        return self.expert_parallel > 1

    @property
    def world_size(self) -> int:
        return (
            self.pipeline_parallel
            * self.data_parallel_replicate
            * self.data_parallel_shard
            * self.context_parallel_replicate
            * self.context_parallel_shard
            * self.tensor_parallel
        )

    @property
    def is_distributed(self) -> bool:
        return self.world_size > 1 or self.has_expert_parallel

    @model_validator(mode="after")
    def _validate(self) -> Self:
        for name, value in self.model_dump().items():
            if value < 1:
                raise ValueError(f"{name} must be >= 1, got {value}")
        pool = self.world_size // self.pipeline_parallel
        if pool % self.expert_parallel != 0:
            raise ValueError(
                f"Total data/context/tensor parallelism degree ({pool}) must be divisible by "
                f"total expert parallelism degree ({self.expert_parallel})."
            )
        return self

    def build(self, log_level: int = logging.INFO) -> "DistributedContext":
        from .configured import DistributedContext

        return DistributedContext(self, log_level)
