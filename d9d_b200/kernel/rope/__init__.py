"""Rotary embedding application as one fused kernel (the reference does this in eager PyTorch:
``d9d/module/block/positional/rope.py:137-173``).  cos/sin are gathered from the cache by ``position_ids`` inside
the kernel; the backward is the same kernel with the rotation transposed."""

from __future__ import annotations

from typing import Any

import torch
from torch.autograd import Function

from .._native import native_ops, on_gpu

STYLE_HALF = 0
STYLE_INTERLEAVED = 1


def rotate_reference(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, style: int) -> torch.Tensor:
    """x: [..., H, D]; cos/sin: [..., rope_dim] broadcast over heads. fp32 oracle."""
    rope_dim = cos.shape[-1]
    xf = x.float()
    xr, xpass = xf[..., :rope_dim], xf[..., rope_dim:]
    c, s = cos.float().unsqueeze(-2), sin.float().unsqueeze(-2)
    if style == STYLE_HALF:
        half = rope_dim // 2
        rot = torch.cat([-xr[..., half:], xr[..., :half]], dim=-1)
    else:
        rot = torch.stack([-xr[..., 1::2], xr[..., 0::2]], dim=-1).flatten(-2)
    return torch.cat([xr * c + rot * s, xpass], dim=-1).to(x.dtype)


class RopeFunction(Function):
    @staticmethod
    def forward(ctx: Any, x: torch.Tensor, cos_cache: torch.Tensor, sin_cache: torch.Tensor, pos: torch.Tensor, style: int):
        shape = x.shape
        x3 = x.reshape(-1, shape[-2], shape[-1])
        if x3.stride(2) != 1 or x3.stride(1) != shape[-1]:
            x3 = x3.contiguous()
        pos_flat = pos.reshape(-1).contiguous()
        ctx.save_for_backward(cos_cache, sin_cache, pos_flat)
        ctx.style = style
        return native_ops().rope_apply(x3, cos_cache, sin_cache, pos_flat, style, False).view(shape)

    @staticmethod
    def backward(ctx: Any, grad_output: torch.Tensor):  # type: ignore[override]
        cos_cache, sin_cache, pos_flat = ctx.saved_tensors
        shape = grad_output.shape
        g3 = grad_output.reshape(-1, shape[-2], shape[-1]).contiguous()
        dx = native_ops().rope_apply(g3, cos_cache, sin_cache, pos_flat, ctx.style, True).view(shape)
        return dx, None, None, None, None


def apply_rope(x: torch.Tensor, cos_cache: torch.Tensor, sin_cache: torch.Tensor, position_ids: torch.Tensor,
               style: int = STYLE_HALF) -> torch.Tensor:
    """Rotate ``x [B,S,H,D]`` by the cached angles of ``position_ids [B,S]``; ``cos/sin_cache`` are ``[max_pos, rope_dim]``."""
    if on_gpu(x) and x.dtype == torch.bfloat16:
        return RopeFunction.apply(x, cos_cache.float(), sin_cache.float(), position_ids, style)
    return rotate_reference(x, cos_cache[position_ids], sin_cache[position_ids], style)


__all__ = ["STYLE_HALF", "STYLE_INTERLEAVED", "apply_rope", "rotate_reference"]
