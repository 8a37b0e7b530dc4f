from __future__ import annotations

import torch

# kernels launched per native op call (used for the bench's ``gpu_launches`` accounting)
_KERNELS_PER_OP = {
    "rms_norm_bwd": 2,
    "moe_build_layout": 3,
    "moe_permute": 2,
    "ce_forward": 3,  # zero-fill of the target-logit buffer + GEMM + finalize
}


class _CountingOps:
    """Thin proxy over ``torch.ops.d9d_b200`` that counts kernel launches of the native extension."""

    def __init__(self, ns):
        self._ns = ns
        self.launches = 0
        self._cache: dict[str, object] = {}

    def __getattr__(self, name: str):
        fn = self._cache.get(name)
        if fn is None:
            op = getattr(self._ns, name)
            n = _KERNELS_PER_OP.get(name, 1)

            def call(*args, __op=op, __n=n, **kwargs):
                self.launches += __n
                return __op(*args, **kwargs)

            fn = call
            self._cache[name] = fn
        return fn


_proxy: _CountingOps | None = None


def native_ops() -> _CountingOps:
    """The native op namespace (loads / builds the extension on first use)."""
    global _proxy
    if _proxy is None:
        from d9d_b200 import ops

        _proxy = _CountingOps(ops.load())
    return _proxy


def native_launch_count() -> int:
    return _proxy.launches if _proxy is not None else 0


def on_gpu(*tensors: torch.Tensor) -> bool:
    """True if the op must run on the native CUDA path."""
    return any(t is not None and t.is_cuda for t in tensors)


def grad_dtype_of(t: torch.Tensor) -> torch.dtype:
    """Dtype autograd expects for the gradient of ``t`` (``grad_dtype`` is only defined on leaves)."""
    try:
        gd = t.grad_dtype if t.is_leaf else None
    except (RuntimeError, AttributeError):
        gd = None
    if gd not in (torch.bfloat16, torch.float32):
        return t.dtype
    return gd
