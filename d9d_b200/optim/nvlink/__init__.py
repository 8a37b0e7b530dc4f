"""Optimizers whose data-parallel gradient reduction and parameter broadcast run in our own NVLink kernels."""

from .sharded_adamw import NvlinkShardedAdamW

__all__ = ["NvlinkShardedAdamW"]
