from __future__ import annotations

import torch


def native_ops():
    """The ``torch.ops.d9d_b200`` namespace (loads / builds the extension on first use)."""
    from d9d_b200 import ops

    return ops.load()


def on_gpu(*tensors: torch.Tensor) -> bool:
    """True if the op must run on the native CUDA path."""
    return any(t is not None and t.is_cuda for t in tensors)
