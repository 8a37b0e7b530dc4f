"""Dataset utilities: data-parallel sharding, length-bucketed shuffling, padding, pooling masks
(reference ``d9d/dataset``) plus a synthetic token dataset for benchmarks."""

from .buffer_sorted import BufferSortedDataset, DatasetImplementingSortKeyProtocol
from .context_parallel import (
    context_parallel_rank_and_size,
    shard_batch_along_sequence,
    shard_batch_for_context_parallel,
    shard_batch_for_sequence_parallel,
)
from .padding import PaddingSide1D, pad_stack_1d
from .pooling import TokenPoolingType, token_pooling_mask_from_attention_mask
from .sharded import ShardedDataset, ShardIndexingMode, shard_dataset_data_parallel
from .synthetic import SyntheticTokenDataset

__all__ = [
    "BufferSortedDataset",
    "DatasetImplementingSortKeyProtocol",
    "PaddingSide1D",
    "ShardIndexingMode",
    "ShardedDataset",
    "SyntheticTokenDataset",
    "TokenPoolingType",
    "context_parallel_rank_and_size",
    "pad_stack_1d",
    "shard_batch_along_sequence",
    "shard_batch_for_context_parallel",
    "shard_batch_for_sequence_parallel",
    "shard_dataset_data_parallel",
    "token_pooling_mask_from_attention_mask",
]
