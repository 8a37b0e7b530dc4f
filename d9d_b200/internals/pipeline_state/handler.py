from __future__ import annotations

from typing import Any

from d9d_b200.core.sharding import ShardingSpecLeaf

from .api import PipelineState
from .storage import PipelineStateStorage


class _GlobalView(PipelineState):
    def __init__(self, storage: PipelineStateStorage):
        self._s = storage

    def __setitem__(self, key: str, value: Any) -> None:
        self._s.store_global((key,), value)

    def __getitem__(self, item: str) -> Any:
        return self._s.acquire_global((item,))

    def __contains__(self, item: str) -> bool:
        return self._s.contains((item,))


class _ShardView(PipelineState):
    def __init__(self, storage: PipelineStateStorage, shard: int):
        self._s, self._shard = storage, shard

    def __setitem__(self, key: str, value: Any) -> None:
        self._s.store_shard((key,), value, self._shard)

    def __getitem__(self, item: str) -> Any:
        return self._s.acquire_shard((item,), self._shard)

    def __contains__(self, item: str) -> bool:
        return self._s.contains((item,))


class PipelineStateHandler:
    """Owns the storage of one step and hands out the global view and per-microbatch views."""

    def __init__(self, sharding_spec: dict[str, ShardingSpecLeaf], num_shards: int):
        self._storage = PipelineStateStorage({(k,): v for k, v in sharding_spec.items()}, num_shards)

    def global_state(self) -> PipelineState:
        return _GlobalView(self._storage)

    def sharded_state(self, shard_id: int) -> PipelineState:
        return _ShardView(self._storage, shard_id)

    def reset(self) -> None:
        self._storage.reset()
