from __future__ import annotations

import torch
from torch import nn

from d9d_b200.kernel.flash_attn import flash_attn_func


class FlashSdpa(nn.Module):
    """Scaled-dot-product attention over ``[B, S, heads, dim]`` tensors with optional learnable per-head sinks and
    a causal sliding window (reference ``d9d/module/block/attention/sdpa/flash.py:9-89``).

    No head-dim padding is needed here (the reference pads to a multiple of 32 to dodge an FA4 bug).
    """

    def __init__(self, num_sinks: int | None = None, window_size: int | None = None) -> None:
        super().__init__()
        if window_size is not None and window_size < 0:
            raise ValueError("`window_size` must be either `None` or a positive integer value")
        self.sinks = nn.Parameter(torch.zeros(num_sinks)) if num_sinks is not None else None
        self._window_size = window_size

    def forward(self, query_states: torch.Tensor, key_states: torch.Tensor, value_states: torch.Tensor,
                attention_mask: torch.Tensor | None, is_causal: bool, scale: float) -> torch.Tensor:
        del attention_mask  # accepted for interface compatibility, never used (reference behaviour)
        if self._window_size is not None and not is_causal:
            raise ValueError("Sliding window attention requires is_causal=True")
        window = (self._window_size, 0) if self._window_size is not None else (None, None)
        out, _ = flash_attn_func(query_states, key_states, value_states, softmax_scale=scale, causal=is_causal,
                                 window_size=window, learnable_sink=self.sinks)
        return out
