from __future__ import annotations

import torch

from ..._native import native_ops, on_gpu
from .._rng import sr_round_reference


def fp32_to_bf16_kernel(val_fp32: torch.Tensor, seed: int) -> torch.Tensor:
    """Unbiased stochastic rounding of an fp32 tensor to a new bf16 tensor.

    16 random bits are added below the kept mantissa and the sum is truncated, so ``E[result] == val_fp32``.  The random
    stream is a pure function of ``(seed, element offset)`` - independent of the launch geometry, reproducible.
    """
    if val_fp32.dtype != torch.float32:
        raise ValueError("Source must be Float32.")
    if not on_gpu(val_fp32):
        return sr_round_reference(val_fp32, seed)
    out = torch.empty_like(val_fp32, dtype=torch.bfloat16, memory_format=torch.contiguous_format)
    native_ops().sr_copy_(out, val_fp32.contiguous(), int(seed))
    return out
