"""Dataset utilities: data-parallel sharding, length-bucketed shuffling, padding, pooling masks
(reference ``d9d/dataset``) plus a synthetic token dataset for benchmarks."""

from .buffer_sorted import BufferSortedDataset, DatasetImplementingSortKeyProtocol
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
    "pad_stack_1d",
    "shard_dataset_data_parallel",
    "token_pooling_mask_from_attention_mask",
]
