"""Ulysses-style context parallelism: swap the sequence split for a head split around the attention kernel."""

from __future__ import annotations

from collections.abc import Callable
from typing import Any

import torch
import torch.distributed as dist
from torch.autograd import Function

AttentionFn = Callable[[torch.Tensor, torch.Tensor, torch.Tensor], torch.Tensor]


class _SeqToHeads(Function):
    """``[B, S/W, H, D]`` (all heads, my tokens) -> ``[B, S, H/W, D]`` (my heads, all tokens, rank-major token order)."""

    @staticmethod
    def forward(ctx: Any, x: torch.Tensor, group: dist.ProcessGroup) -> torch.Tensor:
        ctx.group = group
        return _seq_to_heads(x, group)

    @staticmethod
    def backward(ctx: Any, grad: torch.Tensor):  # type: ignore[override]
        return _heads_to_seq(grad, ctx.group), None


class _HeadsToSeq(Function):
    @staticmethod
    def forward(ctx: Any, x: torch.Tensor, group: dist.ProcessGroup) -> torch.Tensor:
        ctx.group = group
        return _heads_to_seq(x, group)

    @staticmethod
    def backward(ctx: Any, grad: torch.Tensor):  # type: ignore[override]
        return _seq_to_heads(grad, ctx.group), None


def _seq_to_heads(x: torch.Tensor, group: dist.ProcessGroup) -> torch.Tensor:
    world = group.size()
    b, s_local, h, d = x.shape
    # destination-major: piece w = heads [w * h/W, (w+1) * h/W) of my tokens
    send = x.view(b, s_local, world, h // world, d).permute(2, 0, 1, 3, 4).contiguous()
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    # recv[w] = my heads of rank w's tokens -> concatenate the token blocks in rank order
    return recv.permute(1, 0, 2, 3, 4).reshape(b, world * s_local, h // world, d)


def _heads_to_seq(x: torch.Tensor, group: dist.ProcessGroup) -> torch.Tensor:
    world = group.size()
    b, s, h_local, d = x.shape
    s_local = s // world
    send = x.view(b, world, s_local, h_local, d).permute(1, 0, 2, 3, 4).contiguous()  # piece w = rank w's tokens
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    # recv[w] = head block w of my tokens
    return recv.permute(1, 2, 0, 3, 4).reshape(b, s_local, world * h_local, d)


def ulysses_supported(num_heads: int, num_kv_heads: int, world: int) -> bool:
    return num_heads % world == 0 and num_kv_heads % world == 0


def ulysses_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, group: dist.ProcessGroup, attention: AttentionFn,
                      positions: torch.Tensor | None = None) -> torch.Tensor:
    """Run ``attention(q, k, v)`` over the *full* sequence on ``heads / world`` heads per rank.

    ``q [B, S_local, H, D]``, ``k / v [B, S_local, Hk, D]`` with ``H`` and ``Hk`` divisible by the group size.  After the
    exchange tokens are ordered rank-major; when the ranks hold non-contiguous token sets (zig-zag layout) pass the
    concatenated global ``positions [S]`` of that order: the tensors are sorted by position for the kernel (which assumes
    natural order for causal masking) and the result is scattered back.
    """
    world = group.size()
    if not ulysses_supported(q.shape[2], k.shape[2], world):
        raise ValueError(f"Ulysses context parallelism needs head counts divisible by {world}, got {q.shape[2]} / {k.shape[2]}")
    qh, kh, vh = (_SeqToHeads.apply(t, group) for t in (q, k, v))
    if positions is not None:
        order = torch.argsort(positions)
        qh, kh, vh = (t.index_select(1, order) for t in (qh, kh, vh))
    out = attention(qh, kh, vh)
    if positions is not None:
        inverse = torch.empty_like(order)
        inverse[order] = torch.arange(order.numel(), device=order.device)
        out = out.index_select(1, inverse)
    return _HeadsToSeq.apply(out, group)
