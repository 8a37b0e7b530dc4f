"""Gated delta rule.

Per head with state ``S [dk, dv]`` (``q``/``k`` optionally L2-normalised, ``q`` scaled by ``dk^-1/2``):

    S   <- exp(g_t) * S                          # data-dependent decay
    S   <- S + k_t (x) (beta_t * (v_t - S^T k_t))  # delta-rule write
    o_t  = S^T q_t

``recurrent_gated_delta_rule`` is the literal recurrence (oracle); ``chunk_gated_delta_rule`` evaluates it chunk by
chunk with the UT transform so that everything inside a chunk is matmuls (tensor-core shaped work) and only one
``[dk, dv]`` state per head is carried between chunks.
"""

from __future__ import annotations

import torch
import torch.nn.functional as F


def _prep(q: torch.Tensor, k: torch.Tensor, use_qk_l2norm: bool, scale: float | None):
    if use_qk_l2norm:
        # x * rsqrt(sum(x^2) + eps), the convention of the fla kernels / transformers (NOT x / max(|x|, eps): the two differ
        # for the small q / k of freshly initialised models)
        q, k = q.float(), k.float()
        q = q * torch.rsqrt(q.square().sum(dim=-1, keepdim=True) + 1e-6)
        k = k * torch.rsqrt(k.square().sum(dim=-1, keepdim=True) + 1e-6)
    scale = q.shape[-1] ** -0.5 if scale is None else scale
    return q.float() * scale, k.float()


def recurrent_gated_delta_rule(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, g: torch.Tensor, beta: torch.Tensor,
                               use_qk_l2norm: bool = True, scale: float | None = None) -> torch.Tensor:
    """``q, k [B,S,H,dk]``, ``v [B,S,H,dv]``, ``g, beta [B,S,H]`` -> ``o [B,S,H,dv]``."""
    out_dtype = v.dtype
    q, k = _prep(q, k, use_qk_l2norm, scale)
    v, g, beta = v.float(), g.float(), beta.float()
    b, s, h, dk = q.shape
    state = q.new_zeros(b, h, dk, v.shape[-1])
    outs = []
    for t in range(s):
        state = state * g[:, t].exp()[..., None, None]
        pred = torch.einsum("bhkv,bhk->bhv", state, k[:, t])
        state = state + torch.einsum("bhk,bhv->bhkv", k[:, t], beta[:, t][..., None] * (v[:, t] - pred))
        outs.append(torch.einsum("bhkv,bhk->bhv", state, q[:, t]))
    return torch.stack(outs, dim=1).to(out_dtype)


def chunk_gated_delta_rule(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, g: torch.Tensor, beta: torch.Tensor,
                           use_qk_l2norm: bool = True, scale: float | None = None, chunk_size: int = 64) -> torch.Tensor:
    out_dtype = v.dtype
    q, k = _prep(q, k, use_qk_l2norm, scale)
    b, s, h, dk = q.shape
    dv = v.shape[-1]
    c = chunk_size
    pad = (-s) % c
    # [B, H, N, C, d] chunked views; padding steps have k = 0, beta = 0, g = 0 and therefore leave the state alone
    def chunked(t: torch.Tensor) -> torch.Tensor:
        t = t.float().transpose(1, 2)
        if pad:
            t = F.pad(t, (0, 0, 0, pad)) if t.dim() == 4 else F.pad(t, (0, pad))
        return t.reshape(b, h, -1, c, *t.shape[3:])

    q, k, v, g, beta = chunked(q), chunked(k), chunked(v), chunked(g), chunked(beta)
    n = q.shape[2]
    cum = g.cumsum(-1)  # log decay from the chunk start up to and including step i
    # decay from step j (exclusive) to step i (inclusive), i >= j
    rel = (cum[..., :, None] - cum[..., None, :]).tril().exp().tril()
    k_beta = k * beta[..., None]
    v_beta = v * beta[..., None]
    # UT transform: T = (I - A)^-1 with A = -strict_lower(k_beta k^T * rel).  (I - A) is unit lower triangular, so T is ONE batched
    # triangular solve over all (batch, head, chunk) blocks instead of a C-step forward substitution written in Python
    strict = torch.ones(c, c, dtype=torch.bool, device=q.device).tril(-1)
    a = -(k_beta @ k.transpose(-1, -2) * rel).masked_fill(~strict, 0.0)
    eye = torch.eye(c, device=q.device, dtype=q.dtype)
    t_inv = torch.linalg.solve_triangular(eye - a, eye.expand_as(a), upper=False, unitriangular=True)
    w = t_inv @ (k_beta * cum.exp()[..., None])  # what the incoming state contributes to each write
    u = t_inv @ v_beta

    incl = torch.ones(c, c, dtype=torch.bool, device=q.device).tril()
    qk = (q @ k.transpose(-1, -2) * rel).masked_fill(~incl, 0.0)
    state = q.new_zeros(b, h, dk, dv)
    outs = []
    for i in range(n):
        v_new = u[:, :, i] - w[:, :, i] @ state
        o_inter = (q[:, :, i] * cum[:, :, i].exp()[..., None]) @ state
        outs.append(o_inter + qk[:, :, i] @ v_new)
        total = cum[:, :, i, -1]
        k_tail = k[:, :, i] * (total[..., None] - cum[:, :, i]).exp()[..., None]
        state = state * total.exp()[..., None, None] + k_tail.transpose(-1, -2) @ v_new
    o = torch.stack(outs, dim=2).reshape(b, h, n * c, dv)[:, :, :s]
    return o.transpose(1, 2).contiguous().to(out_dtype)
