"""Attention layers."""

from .grouped_query import GroupedQueryAttention
from .linear import GatedDeltaNet
from .multi_head_latent import LowRankProjection, MultiHeadLatentAttention

__all__ = ["GatedDeltaNet", "GroupedQueryAttention", "LowRankProjection", "MultiHeadLatentAttention"]
