"""CUDA fast path for attention.

* forward-only work (inference, evaluation, ``torch.no_grad``): the hand-written tcgen05 flash-attention forward
  (``ops/csrc/flash_attn_fwd.cu``: TMA-fed Q/K/V tiles, S and O accumulators in TMEM, online softmax in registers,
  P·V straight from shared memory, GQA and causal masking, LSE output);
* training: PyTorch SDPA (cuDNN flash kernels, library code) — the matching backward kernel is still to be written.

``D9D_NATIVE_ATTENTION=0`` forces the SDPA path everywhere.
"""

from __future__ import annotations

import math
import os

import torch
import torch.nn.functional as F

from .._native import native_ops


def native_forward_supported(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor) -> bool:
    return (q.is_cuda and q.dtype == torch.bfloat16 and k.dtype == torch.bfloat16 and v.dtype == torch.bfloat16
            and q.dim() == 4 and q.shape[-1] in (64, 128) and v.shape[-1] == q.shape[-1] and q.shape[2] % k.shape[2] == 0
            and os.environ.get("D9D_NATIVE_ATTENTION", "1") != "0")


def flash_attention_forward(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, softmax_scale: float | None, causal: bool
                            ) -> tuple[torch.Tensor, torch.Tensor]:
    """Forward only: ``(out [B,S,H,D], lse [B,H,S] fp32)`` from the native kernel."""
    scale = softmax_scale if softmax_scale is not None else 1.0 / math.sqrt(q.shape[-1])
    return native_ops().flash_attn_fwd(q.contiguous(), k.contiguous(), v.contiguous(), float(scale), bool(causal))


def flash_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, softmax_scale: float | None, causal: bool) -> torch.Tensor:
    """q: [B,S,H,D]; k/v: [B,S,Hk,D] -> [B,S,H,D]."""
    needs_grad = torch.is_grad_enabled() and (q.requires_grad or k.requires_grad or v.requires_grad)
    if not needs_grad and native_forward_supported(q, k, v):
        return flash_attention_forward(q, k, v, softmax_scale, causal)[0]
    qh, kh, vh = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
    out = F.scaled_dot_product_attention(qh, kh, vh, is_causal=causal, scale=softmax_scale, enable_gqa=qh.shape[1] != kh.shape[1])
    return out.transpose(1, 2)
