"""PyTree micro-batching: split / merge arbitrary trees of tensors and lists into N shards.

Used for pipeline microbatch splitting and for the per-step pipeline state.
Parity: reference ``d9d/core/sharding`` (``shard.py:92-136``, ``unshard.py:53-99``, ``auto_spec.py:26-64``).
"""

from .spec import ShardingSpec, ShardingSpecLeaf, SpecReplicate, SpecShard
from .tree import shard_spec_nothing, shard_spec_on_dim, shard_tree, unshard_tree

__all__ = [
    "ShardingSpec",
    "ShardingSpecLeaf",
    "SpecReplicate",
    "SpecShard",
    "shard_spec_nothing",
    "shard_spec_on_dim",
    "shard_tree",
    "unshard_tree",
]
