"""``gmm(a, b, batch_sizes, trans_a, trans_b)`` with the semantics of nv-grouped-gemm's backend.

* ``trans_a=False``: ``a [T, K]``, ``b [E, K, N]`` (``[E, N, K]`` when ``trans_b``) -> ``[T, N]``; rows of ``a`` are
  grouped by expert according to ``batch_sizes`` (CPU int64).
* ``trans_a=True`` : ``a [T, M]``, ``b [T, N]`` -> ``[E, M, N]`` (per-expert ``a_e^T @ b_e``).
"""

from __future__ import annotations

import torch

_STATE = {"grouped_mm": hasattr(torch, "_grouped_mm")}


def _offsets(batch_sizes: torch.Tensor, device: torch.device) -> torch.Tensor:
    return torch.cumsum(batch_sizes.to(device=device, non_blocking=True), dim=0).to(torch.int32)


def _loop(a, b, batch_sizes, trans_a, trans_b):
    sizes = batch_sizes.tolist()
    if trans_a:
        out = a.new_zeros(len(sizes), a.shape[1], b.shape[1])
        start = 0
        for e, n in enumerate(sizes):
            if n:
                torch.mm(a[start : start + n].t(), b[start : start + n], out=out[e])
            start += n
        return out
    n_out = b.shape[1] if trans_b else b.shape[2]
    out = a.new_empty(a.shape[0], n_out)
    start = 0
    for e, n in enumerate(sizes):
        if n:
            torch.mm(a[start : start + n], b[e].t() if trans_b else b[e], out=out[start : start + n])
        start += n
    return out


def gmm(a: torch.Tensor, b: torch.Tensor, batch_sizes: torch.Tensor, trans_a: bool = False, trans_b: bool = False) -> torch.Tensor:
    if _STATE["grouped_mm"] and a.is_cuda and a.dtype == torch.bfloat16 and a.shape[0] > 0:
        try:
            offs = _offsets(batch_sizes, a.device)
            if trans_a:
                return torch._grouped_mm(a.t(), b, offs=offs)
            return torch._grouped_mm(a, b.transpose(-2, -1) if trans_b else b, offs=offs)
        except (RuntimeError, NotImplementedError):
            _STATE["grouped_mm"] = False
    return _loop(a, b, batch_sizes, trans_a, trans_b)
