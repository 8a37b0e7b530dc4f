"""Per-head q/k RMSNorm fused with rotary embedding (one launch forward, two backward; ``ops/csrc/qk_norm_rope.cu``)."""

from __future__ import annotations

from typing import Any

import torch
from torch.autograd import Function

from .._native import native_ops, on_gpu


class _QkNormRope(Function):
    @staticmethod
    def forward(ctx: Any, q: torch.Tensor, k: torch.Tensor, wq: torch.Tensor, wk: torch.Tensor, cos_t: torch.Tensor,
                sin_t: torch.Tensor, eps: float, zero_centered: bool, style: int):
        q3, k3 = q.reshape(-1, q.shape[-2], q.shape[-1]), k.reshape(-1, k.shape[-2], k.shape[-1])
        qo, ko, inv = native_ops().qk_norm_rope_fwd(q3, k3, wq, wk, cos_t, sin_t, eps, zero_centered, style)
        ctx.save_for_backward(q3, k3, wq, wk, cos_t, sin_t, inv)
        ctx.zero_centered, ctx.style, ctx.q_shape, ctx.k_shape = zero_centered, style, q.shape, k.shape
        return qo.view(q.shape), ko.view(k.shape)

    @staticmethod
    def backward(ctx: Any, dqo: torch.Tensor, dko: torch.Tensor):  # type: ignore[override]
        q3, k3, wq, wk, cos_t, sin_t, inv = ctx.saved_tensors
        dq, dk, dwq, dwk = native_ops().qk_norm_rope_bwd(dqo.reshape(q3.shape).contiguous(), dko.reshape(k3.shape).contiguous(), q3, k3,
                                                         wq, wk, cos_t, sin_t, inv, ctx.zero_centered, ctx.style)
        return dq.view(ctx.q_shape), dk.view(ctx.k_shape), dwq.to(wq.dtype), dwk.to(wk.dtype), None, None, None, None, None


def fused_qk_norm_rope_supported(q: torch.Tensor, k: torch.Tensor, wq: torch.Tensor, wk: torch.Tensor, rope_dim: int) -> bool:
    d = q.shape[-1]
    return (on_gpu(q) and q.dtype == torch.bfloat16 and k.dtype == torch.bfloat16 and wq.dtype == torch.bfloat16
            and wk.dtype == torch.bfloat16 and d in (64, 128, 256) and k.shape[-1] == d and rope_dim <= d
            and rope_dim % 16 == 0)


def qk_norm_rope(q: torch.Tensor, k: torch.Tensor, wq: torch.Tensor, wk: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor,
                 eps: float, zero_centered: bool, style: int) -> tuple[torch.Tensor, torch.Tensor]:
    """``q [..., Hq, D]``, ``k [..., Hk, D]`` (views of the projection outputs are fine), ``cos/sin [..., rope_dim]``
    per token -> normalised and rotated ``q, k`` (contiguous)."""
    rope_dim = cos.shape[-1]
    cos_t = cos.reshape(-1, rope_dim).float().contiguous()
    sin_t = sin.reshape(-1, rope_dim).float().contiguous()
    return _QkNormRope.apply(q, k, wq, wk, cos_t, sin_t, eps, zero_centered, style)
