from __future__ import annotations

import dataclasses

from d9d_b200.core.sharding import ShardingSpec


@dataclasses.dataclass
class PipelineShardingSpec:
    """How pipeline inputs / kwargs are split into microbatches (``None`` => split every tensor on dim 0)."""

    input_data: ShardingSpec | None = None
    input_kwargs: ShardingSpec | None = None
