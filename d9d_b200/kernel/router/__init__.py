"""Fused MoE routing: fp32 softmax over experts + top-k (+ selection bias, + renormalisation) and its backward."""

from __future__ import annotations

from typing import Any

import torch
import torch.nn.functional as F
from torch.autograd import Function

from .._native import native_ops, on_gpu


def route_topk_reference(logits: torch.Tensor, top_k: int, renormalize: bool, expert_bias: torch.Tensor | None = None):
    """Plain PyTorch definition (CPU path and numerics oracle): returns ``(indices [T,k] int64, probs [T,k] fp32)``."""
    probs = F.softmax(logits, dim=-1, dtype=torch.float32)
    if expert_bias is None:
        chosen_p, chosen = torch.topk(probs, k=top_k, dim=-1)
    else:
        chosen = torch.topk(probs + expert_bias, k=top_k, dim=-1).indices
        chosen_p = probs.gather(-1, chosen)
    if renormalize:
        chosen_p = chosen_p / (chosen_p.sum(dim=-1, keepdim=True) + 1e-20)
    return chosen, chosen_p


class _RouteTopK(Function):
    @staticmethod
    def forward(ctx: Any, logits: torch.Tensor, top_k: int, renormalize: bool, expert_bias: torch.Tensor | None):
        idx, probs = native_ops().router_topk_fwd(logits, expert_bias, top_k, renormalize)
        ctx.save_for_backward(logits, idx)
        ctx.renormalize = renormalize
        ctx.mark_non_differentiable(idx)
        return idx, probs

    @staticmethod
    def backward(ctx: Any, _didx: torch.Tensor, dprobs: torch.Tensor):  # type: ignore[override]
        logits, idx = ctx.saved_tensors
        return native_ops().router_topk_bwd(logits, idx, dprobs.float().contiguous(), ctx.renormalize), None, None, None


def route_topk(logits: torch.Tensor, top_k: int, renormalize: bool, expert_bias: torch.Tensor | None = None):
    """``logits [..., E]`` -> ``(indices [..., k] int64, probabilities [..., k] fp32)``."""
    if on_gpu(logits) and logits.dtype == torch.bfloat16 and logits.shape[-1] <= 1024 and top_k <= 32:
        flat = logits.reshape(-1, logits.shape[-1])
        if not flat.is_contiguous():
            flat = flat.contiguous()
        idx, probs = _RouteTopK.apply(flat, top_k, renormalize, None if expert_bias is None else expert_bias.float().contiguous())
        return idx.view(*logits.shape[:-1], top_k), probs.view(*logits.shape[:-1], top_k)
    return route_topk_reference(logits, top_k, renormalize, expert_bias)


__all__ = ["route_topk", "route_topk_reference"]
