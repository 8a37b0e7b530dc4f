from __future__ import annotations

import copy
from typing import Any

import torch
import torch.utils._pytree as pytree

from d9d_b200.core.sharding import ShardingSpecLeaf, SpecReplicate, SpecShard, shard_tree, unshard_tree

StateKey = tuple[str, ...]


def _detached(tree: Any) -> Any:
    return pytree.tree_map(lambda x: x.detach() if isinstance(x, torch.Tensor) else x, tree)


class _PerShard(dict):
    """Marker container: shard index -> value."""


class PipelineStateStorage:
    """Stores each key either as one global value or as per-shard values and converts lazily between the two
    using a sharding spec that is given up front or inferred at first write:

    * written globally: 0-d tensors / non-tensors replicate, tensors and lists split on dim 0;
    * written per shard: 0-d tensors and python scalars *stack* into a new dim 0, tensors / lists concatenate.
    """

    def __init__(self, sharding_spec: dict[StateKey, ShardingSpecLeaf], num_shards: int):
        self._given = copy.deepcopy(sharding_spec)
        self._num_shards = num_shards
        self._values: dict[StateKey, Any] = {}
        self._specs: dict[StateKey, ShardingSpecLeaf] = {}

    def _infer_from_global(self, key: StateKey, value: Any) -> ShardingSpecLeaf:
        if key in self._given:
            return self._given[key]
        if isinstance(value, torch.Tensor):
            return SpecReplicate() if value.ndim == 0 else SpecShard(0)
        if isinstance(value, list):
            return SpecShard(0)
        return SpecReplicate()

    def _infer_from_shard(self, key: StateKey, value: Any) -> ShardingSpecLeaf:
        if key in self._given:
            return self._given[key]
        if isinstance(value, torch.Tensor):
            return SpecShard(0, do_stack=value.ndim == 0)
        if isinstance(value, list):
            return SpecShard(0)
        return SpecShard(0, do_stack=True)

    def store_global(self, key: StateKey, value: Any) -> None:
        value = _detached(value)
        self._specs.setdefault(key, self._infer_from_global(key, value))
        self._values[key] = value

    def store_shard(self, key: StateKey, value: Any, shard_id: int) -> None:
        slot = self._values.setdefault(key, _PerShard())
        if not isinstance(slot, _PerShard):
            raise ValueError(f"Trying to store sharded state into an unsharded one: {key}")
        value = _detached(value)
        self._specs.setdefault(key, self._infer_from_shard(key, value))
        slot[shard_id] = value

    def _require(self, key: StateKey) -> Any:
        if key not in self._values:
            raise ValueError(f"Cannot access non-existing state {key}")
        return self._values[key]

    def acquire_global(self, key: StateKey) -> Any:
        value = self._require(key)
        if isinstance(value, _PerShard):
            value = unshard_tree([value[i] for i in range(self._num_shards)], self._specs[key])
            self._values[key] = value
        return value

    def acquire_shard(self, key: StateKey, shard: int) -> Any:
        value = self._require(key)
        if not isinstance(value, _PerShard):
            pieces = shard_tree(value, self._specs[key], num_shards=self._num_shards, enforce_even_split=True)
            value = _PerShard(enumerate(pieces))
            self._values[key] = value
        return value[shard]

    def contains(self, key: StateKey) -> bool:
        return key in self._values

    def reset(self) -> None:
        self._values.clear()
