from __future__ import annotations

import torch

_SIGNED = {8: torch.int8, 16: torch.int16, 32: torch.int32, 64: torch.int64}
_UNSIGNED = {8: torch.uint8, 16: torch.uint16, 32: torch.uint32, 64: torch.uint64}


def get_int_dtype(bitwidth: int, signed: bool) -> torch.dtype:
    """Integer dtype of the given width - what a kernel bit-casts a float of that width to.

    Reference ``d9d/kernel/general/get_int_dtype.py:5-7`` resolves a Triton dtype; the CUDA kernels here do their
    bit-casts in C++, so the host-side helper answers in torch dtypes (used by tests and by the CPU oracles).
    """
    table = _SIGNED if signed else _UNSIGNED
    if bitwidth not in table:
        raise ValueError(f"Unsupported integer bit width: {bitwidth}")
    return table[bitwidth]
