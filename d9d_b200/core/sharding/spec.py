from __future__ import annotations

import dataclasses

from d9d_b200.core.types import PyTree


@dataclasses.dataclass(frozen=True, slots=True)
class SpecReplicate:
    """Leaf is shared (by reference) between all shards."""


@dataclasses.dataclass(frozen=True, slots=True)
class SpecShard:
    """Leaf is split along ``dim``.

    With ``do_stack`` the dimension must have exactly ``num_shards`` entries and is squeezed away on split /
    re-created by stacking on merge.
    """

    dim: int
    do_stack: bool = False


ShardingSpecLeaf = SpecReplicate | SpecShard
ShardingSpec = PyTree[ShardingSpecLeaf]
