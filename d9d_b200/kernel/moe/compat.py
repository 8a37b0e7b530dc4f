"""Reference-compatible MoE routing helpers, expressed with device-side torch ops (no host sync).

These exist for API parity with ``d9d.kernel.moe``; the model path uses :mod:`.layout`.
"""

from __future__ import annotations

import torch


def fused_indices_to_multihot(indices: torch.Tensor, probs_indices: torch.Tensor, num_of_local_experts: int):
    """``(indices[T,k], probs[T,k]) -> (routing_map bool[T,E], probs[T,E])``; ids outside ``[0,E)`` are dropped.

    Reference: ``d9d/kernel/moe/indices_to_multihot.py:157-268``.
    """
    T = indices.shape[0]
    valid = (indices >= 0) & (indices < num_of_local_experts)
    # dropped ids are redirected to column 0 with a zero contribution; scatter_add makes that a no-op (a plain scatter
    # could overwrite a genuine hit on column 0)
    safe = torch.where(valid, indices, torch.zeros_like(indices)).long()
    hits = torch.zeros(T, num_of_local_experts, dtype=torch.int32, device=indices.device)
    hits.scatter_add_(1, safe, valid.int())
    probs = torch.zeros(T, num_of_local_experts, dtype=probs_indices.dtype, device=indices.device)
    probs.scatter_add_(1, safe, probs_indices * valid.to(probs_indices.dtype))
    return hits > 0, probs


def moe_permute_with_probs(inp: torch.Tensor, probs: torch.Tensor, routing_map: torch.Tensor, num_out_tokens: int = -1):
    """Group tokens by expert (expert-major, token order preserved inside an expert).

    Returns ``(permuted [R,H], permuted_probs [R], row_id_map)`` where ``row_id_map`` is the ``[R]`` source-token
    index of every output row plus the bookkeeping needed by :func:`moe_unpermute_mask`.
    Reference: ``d9d/kernel/moe/permute_with_probs.py:633-741`` (row-id-map encoding differs; it is opaque).
    """
    T, E = routing_map.shape
    mask_t = routing_map.bool().t().contiguous()  # [E, T] expert-major
    if num_out_tokens is not None and num_out_tokens >= 0:
        flat_idx = mask_t.reshape(-1).nonzero(as_tuple=False).flatten()[:num_out_tokens]
    else:
        flat_idx = mask_t.reshape(-1).nonzero(as_tuple=False).flatten()
    token_idx = flat_idx % T
    expert_idx = flat_idx // T
    permuted = inp.index_select(0, token_idx)
    permuted_probs = probs[token_idx, expert_idx] if probs is not None else None
    row_id_map = torch.stack([token_idx, expert_idx], dim=0)
    return permuted, permuted_probs, row_id_map


def moe_unpermute_mask(inp: torch.Tensor, row_id_map: torch.Tensor, merging_probs: torch.Tensor | None = None,
                       restore_shape: torch.Size | None = None) -> torch.Tensor:
    """Inverse of :func:`moe_permute_with_probs`: accumulate (fp32) the rows of every token, optionally weighted."""
    token_idx, expert_idx = row_id_map[0], row_id_map[1]
    if restore_shape is None:
        raise ValueError("restore_shape is required")
    out = torch.zeros(restore_shape, dtype=torch.float32, device=inp.device)
    vals = inp.float()
    if merging_probs is not None:
        vals = vals * merging_probs[token_idx, expert_idx].float().unsqueeze(-1)
    out = out.index_add(0, token_idx, vals)
    return out.to(inp.dtype)
