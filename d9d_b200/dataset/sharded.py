from __future__ import annotations

import enum
from collections.abc import Sized
from typing import Any, TypeVar

from torch.distributed.checkpoint.stateful import Stateful
from torch.utils.data import Dataset

from d9d_b200.core.dist_context import BATCH_DOMAIN, DistributedContext

_T_co = TypeVar("_T_co", covariant=True)


class ShardIndexingMode(enum.StrEnum):
    sequential = "sequential"  # round robin: shard r sees r, r+N, r+2N, ...
    chunked = "chunked"  # contiguous blocks of ceil(len/N)


class ShardedDataset(Dataset[_T_co], Stateful):
    """View of one shard of a dataset (data parallelism).  With ``pad_to_equal_size_across_shards`` every shard
    reports the ceiling length and out-of-range indices repeat the last element, so collectives never hang on an
    uneven tail.  Parity: reference ``d9d/dataset/sharded.py:39-201``.
    """

    def __init__(self, dataset: Dataset[_T_co], total_shards: int, current_shard: int, indexing_mode: ShardIndexingMode,
                 pad_to_equal_size_across_shards: bool):
        if not isinstance(dataset, Sized):
            raise ValueError("Dataset should implement __len__ method")
        if indexing_mode not in (ShardIndexingMode.sequential, ShardIndexingMode.chunked):
            raise ValueError(f"Unknown shard indexing mode: {indexing_mode}")
        self._dataset = dataset
        self._n = total_shards
        self._rank = current_shard
        self._mode = indexing_mode
        self._pad = pad_to_equal_size_across_shards

    def _ceil_len(self) -> int:
        return -(-len(self._dataset) // self._n)

    def __len__(self) -> int:
        total = len(self._dataset)
        if self._pad:
            return self._ceil_len()
        if self._mode == ShardIndexingMode.sequential:
            return total // self._n + (1 if self._rank < total % self._n else 0)
        start = self._ceil_len() * self._rank
        return max(0, min(self._ceil_len(), total - start))

    def __getitem__(self, index: int) -> _T_co:
        if self._mode == ShardIndexingMode.sequential:
            base = index * self._n + self._rank
        else:
            base = self._ceil_len() * self._rank + index
        return self._dataset[min(base, len(self._dataset) - 1)]

    def state_dict(self) -> dict[str, Any]:
        state: dict[str, Any] = {"total_shards": self._n, "current_shard": self._rank}
        if isinstance(self._dataset, Stateful):
            state["dataset"] = self._dataset.state_dict()
        return state

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:
        if state_dict["total_shards"] != self._n:
            raise ValueError("Shard count mismatch")
        if state_dict["current_shard"] != self._rank:
            raise ValueError("Shard index mismatch: a dataset shard must be restored on the rank that saved it")
        if isinstance(self._dataset, Stateful):
            self._dataset.load_state_dict(state_dict["dataset"])


def shard_dataset_data_parallel(dataset: Dataset[_T_co], dist_context: DistributedContext,
                                indexing_mode: ShardIndexingMode = ShardIndexingMode.sequential,
                                pad_to_equal_size_across_shards: bool = True) -> Dataset[_T_co]:
    """Shard over the ``dp`` dimension of the batch mesh (identity when not distributed)."""
    if dist_context.mesh_params.is_distributed:
        dp = dist_context.mesh_for(BATCH_DOMAIN)["dp"]
        n, r = dp.size(), dp.get_local_rank()
    else:
        n, r = 1, 0
    return ShardedDataset(dataset, total_shards=n, current_shard=r, indexing_mode=indexing_mode,
                          pad_to_equal_size_across_shards=pad_to_equal_size_across_shards)
