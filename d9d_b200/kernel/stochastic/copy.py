from __future__ import annotations

import torch

from .._native import native_ops, on_gpu
from ._rng import draw_seed, sr_round_reference


def copy_fp32_to_bf16_stochastic_(
    target: torch.Tensor, source: torch.Tensor, generator: torch.Generator | None = None
) -> torch.Tensor:
    """In-place ``target(bf16) <- SR(source(fp32))``. Reference: ``d9d/kernel/stochastic/copy.py:34-85``."""
    if target.dtype != torch.bfloat16:
        raise ValueError("Target must be BFloat16.")
    if source.dtype != torch.float32:
        raise ValueError("Source must be Float32.")
    if target.shape != source.shape:
        raise ValueError("Shape mismatch between target and source.")
    if not target.is_contiguous():
        raise ValueError("Target must be contiguous since it is an in-place kernel.")
    seed = draw_seed(generator)
    if on_gpu(target):
        native_ops().sr_copy_(target, source.contiguous(), seed)
    else:
        target.copy_(sr_round_reference(source, seed))
    return target
