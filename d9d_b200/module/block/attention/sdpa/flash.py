from __future__ import annotations

import enum

import torch
import torch.distributed as dist
from torch import nn

from d9d_b200.kernel.context_parallel import (
    ContextParallelLayout,
    local_sequence_indices,
    ring_attention,
    ulysses_attention,
    ulysses_supported,
)
from d9d_b200.kernel.flash_attn import flash_attn_func


class ContextParallelMode(enum.StrEnum):
    """How a context-parallel group exchanges data around attention (see ``d9d_b200.kernel.context_parallel``)."""

    auto = "auto"  # Ulysses when the head counts are divisible by the group size, ring otherwise
    ulysses = "ulysses"
    ring = "ring"


class FlashSdpa(nn.Module):
    """Scaled-dot-product attention over ``[B, S, heads, dim]`` tensors with optional learnable per-head sinks and
    a causal sliding window (reference ``d9d/module/block/attention/sdpa/flash.py:9-89``).

    No head-dim padding is needed here (the reference pads to a multiple of 32 to dodge an FA4 bug).

    With :meth:`enable_context_parallel` the tensors passed to ``forward`` hold only this rank's tokens of every sequence
    and attention is computed over the tokens of the whole group (net-new relative to the reference).
    """

    def __init__(self, num_sinks: int | None = None, window_size: int | None = None) -> None:
        super().__init__()
        if window_size is not None and window_size < 0:
            raise ValueError("`window_size` must be either `None` or a positive integer value")
        self.sinks = nn.Parameter(torch.zeros(num_sinks)) if num_sinks is not None else None
        self._window_size = window_size
        self._cp_group: dist.ProcessGroup | None = None
        self._cp_mode = ContextParallelMode.auto
        self._cp_layout = ContextParallelLayout.zigzag
        self._cp_mask_cache: dict = {}

    def enable_context_parallel(self, group: dist.ProcessGroup, mode: ContextParallelMode = ContextParallelMode.auto,
                                layout: ContextParallelLayout = ContextParallelLayout.zigzag) -> None:
        """Attend over the tokens of every rank of ``group``; ``layout`` says which tokens each rank holds (it must match
        how the batch was sharded, see ``d9d_b200.dataset.shard_batch_for_context_parallel``)."""
        self._cp_group = group if group.size() > 1 else None
        self._cp_mode, self._cp_layout = ContextParallelMode(mode), ContextParallelLayout(layout)
        self._cp_mask_cache = {}

    def _local_attention(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, is_causal: bool, scale: float,
                         sinks: torch.Tensor | None) -> torch.Tensor:
        window = (self._window_size, 0) if self._window_size is not None else (None, None)
        out, _ = flash_attn_func(q, k, v, softmax_scale=scale, causal=is_causal, window_size=window, learnable_sink=sinks)
        return out

    def _context_parallel_attention(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, is_causal: bool, scale: float
                                    ) -> torch.Tensor:
        group = self._cp_group
        assert group is not None
        world, rank = group.size(), group.rank()
        seq = q.shape[1] * world
        positions = torch.stack([local_sequence_indices(seq, world, r, self._cp_layout) for r in range(world)])
        mode = self._cp_mode
        if mode == ContextParallelMode.auto:
            mode = ContextParallelMode.ulysses if ulysses_supported(q.shape[2], k.shape[2], world) else ContextParallelMode.ring
        if mode == ContextParallelMode.ulysses:
            heads_local = q.shape[2] // world
            sinks = self.sinks[rank * heads_local : (rank + 1) * heads_local] if self.sinks is not None else None
            contiguous = self._cp_layout == ContextParallelLayout.contiguous
            return ulysses_attention(q, k, v, group, lambda a, b, c: self._local_attention(a, b, c, is_causal, scale, sinks),
                                     positions=None if contiguous else positions.reshape(-1).to(q.device))
        if self.sinks is not None or self._window_size is not None:
            raise ValueError("ring context parallelism does not support attention sinks / sliding windows; use the ulysses mode")
        return ring_attention(q, k, v, group, positions, softmax_scale=scale, causal=is_causal, mask_cache=self._cp_mask_cache)

    def forward(self, query_states: torch.Tensor, key_states: torch.Tensor, value_states: torch.Tensor,
                attention_mask: torch.Tensor | None, is_causal: bool, scale: float) -> torch.Tensor:
        del attention_mask  # accepted for interface compatibility, never used (reference behaviour)
        if self._window_size is not None and not is_causal:
            raise ValueError("Sliding window attention requires is_causal=True")
        if self._cp_group is not None:
            return self._context_parallel_attention(query_states, key_states, value_states, is_causal, scale)
        return self._local_attention(query_states, key_states, value_states, is_causal, scale, self.sinks)
