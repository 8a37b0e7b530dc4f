"""Attention entry points with the reference's ``flash_attn_func`` / ``flash_attn_varlen_func`` signatures
(``d9d/kernel/flash_attn/function.py:313-436``): causal, sliding window, GQA/MQA, softcap, learnable per-head
sinks (with analytic ``dsink``), optional LSE.

Backends:
* ``attention_reference``: exact fp32 math (CPU, and features the fast path does not cover yet),
* CUDA fast path: fused flash kernel (no [S,S] materialisation).
"""

from __future__ import annotations

import math

import torch


def _window_mask(sq: int, sk: int, causal: bool, window: tuple[int | None, int | None], device) -> torch.Tensor | None:
    left, right = window
    if not causal and left is None and right is None:
        return None
    # bottom-right aligned relative position (FlashAttention convention)
    qi = torch.arange(sq, device=device)[:, None] + (sk - sq)
    kj = torch.arange(sk, device=device)[None, :]
    keep = torch.ones(sq, sk, dtype=torch.bool, device=device)
    if causal:
        keep &= kj <= qi
    if left is not None and left >= 0:
        keep &= kj >= qi - left
    if right is not None and right >= 0 and not causal:
        keep &= kj <= qi + right
    return keep


def attention_reference(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    softmax_scale: float | None = None,
    causal: bool = False,
    window_size: tuple[int | None, int | None] = (None, None),
    learnable_sink: torch.Tensor | None = None,
    softcap: float = 0.0,
) -> tuple[torch.Tensor, torch.Tensor]:
    """fp32 attention oracle. q: [B,Sq,H,D]; k/v: [B,Sk,Hk,D(v)]. Returns (out [B,Sq,H,Dv], lse [B,H,Sq])."""
    B, Sq, H, D = q.shape
    Sk, Hk = k.shape[1], k.shape[2]
    scale = softmax_scale if softmax_scale is not None else 1.0 / math.sqrt(D)
    group = H // Hk
    qf = q.float().permute(0, 2, 1, 3)  # [B,H,Sq,D]
    kf = k.float().permute(0, 2, 1, 3).repeat_interleave(group, dim=1)
    vf = v.float().permute(0, 2, 1, 3).repeat_interleave(group, dim=1)
    s = torch.matmul(qf, kf.transpose(-1, -2)) * scale
    if softcap and softcap > 0:
        s = torch.tanh(s / softcap) * softcap
    keep = _window_mask(Sq, Sk, causal, window_size, q.device)
    if keep is not None:
        s = s.masked_fill(~keep, float("-inf"))
    if learnable_sink is not None:
        sink = learnable_sink.float().view(1, H, 1, 1).expand(B, H, Sq, 1)
        full = torch.cat([s, sink], dim=-1)
        lse = torch.logsumexp(full, dim=-1)
        p = torch.exp(s - lse.unsqueeze(-1))
    else:
        lse = torch.logsumexp(s, dim=-1)
        p = torch.exp(s - lse.unsqueeze(-1))
    p = torch.nan_to_num(p, nan=0.0)
    out = torch.matmul(p, vf).permute(0, 2, 1, 3).to(q.dtype)
    return out, lse


def _fast_path_ok(q, k, causal, window_size, learnable_sink, softcap) -> bool:
    if not q.is_cuda or q.dtype not in (torch.bfloat16, torch.float16):
        return False
    if learnable_sink is not None or (softcap and softcap > 0):
        return False
    if window_size[0] is not None or window_size[1] is not None:
        return False
    if causal and q.shape[1] != k.shape[1]:
        return False
    return True


def flash_attn_func(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    softmax_scale: float | None = None,
    causal: bool = False,
    window_size: tuple[int | None, int | None] = (None, None),
    learnable_sink: torch.Tensor | None = None,
    softcap: float = 0.0,
    num_splits: int = 1,
    pack_gqa: bool | None = None,
    deterministic: bool = False,
    return_lse: bool = False,
) -> tuple[torch.Tensor, torch.Tensor | None]:
    """Returns ``(output [B,S,H,Dv], lse [B,H,S] or None)``."""
    del num_splits, pack_gqa, deterministic
    if _fast_path_ok(q, k, causal, window_size, learnable_sink, softcap):
        from .native import flash_attention, flash_attention_forward, native_forward_supported

        needs_grad = torch.is_grad_enabled() and (q.requires_grad or k.requires_grad or v.requires_grad)
        if return_lse and not needs_grad and native_forward_supported(q, k, v):
            return flash_attention_forward(q, k, v, softmax_scale, causal)
        if not return_lse:
            return flash_attention(q, k, v, softmax_scale, causal), None
    out, lse = attention_reference(q, k, v, softmax_scale, causal, window_size, learnable_sink, softcap)
    return out, (lse if return_lse else None)


def flash_attn_varlen_func(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    cu_seqlens_q: torch.Tensor | None = None,
    cu_seqlens_k: torch.Tensor | None = None,
    max_seqlen_q: int | None = None,
    max_seqlen_k: int | None = None,
    seqused_q: torch.Tensor | None = None,
    seqused_k: torch.Tensor | None = None,
    page_table: torch.Tensor | None = None,
    softmax_scale: float | None = None,
    causal: bool = False,
    window_size: tuple[int | None, int | None] = (None, None),
    learnable_sink: torch.Tensor | None = None,
    softcap: float = 0.0,
    num_splits: int = 1,
    pack_gqa: bool | None = None,
    deterministic: bool = False,
    return_lse: bool = False,
) -> tuple[torch.Tensor, torch.Tensor | None]:
    """Packed variable-length attention: q ``[total_q, H, D]``, k/v ``[total_k, Hk, D]`` with ``cu_seqlens``."""
    del max_seqlen_q, max_seqlen_k, num_splits, pack_gqa, deterministic
    if page_table is not None or seqused_q is not None or seqused_k is not None:
        raise NotImplementedError("paged KV / seqused are not supported")
    if cu_seqlens_q is None or cu_seqlens_k is None:
        raise ValueError("cu_seqlens_q and cu_seqlens_k are required")
    bq, bk = cu_seqlens_q.tolist(), cu_seqlens_k.tolist()
    outs, lses = [], []
    for i in range(len(bq) - 1):
        o, l = flash_attn_func(
            q[bq[i] : bq[i + 1]].unsqueeze(0), k[bk[i] : bk[i + 1]].unsqueeze(0), v[bk[i] : bk[i + 1]].unsqueeze(0),
            softmax_scale, causal, window_size, learnable_sink, softcap, return_lse=return_lse,
        )
        outs.append(o.squeeze(0))
        if return_lse:
            lses.append(l.squeeze(0))
    out = torch.cat(outs, dim=0)
    return out, (torch.cat(lses, dim=-1) if return_lse else None)
