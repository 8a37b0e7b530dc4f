"""Attention layers."""

from .grouped_query import GroupedQueryAttention
from .multi_head_latent import LowRankProjection, MultiHeadLatentAttention

__all__ = ["GroupedQueryAttention", "LowRankProjection", "MultiHeadLatentAttention"]
