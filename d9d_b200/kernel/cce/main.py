"""Fused linear + cross-entropy without materialising ``[tokens, vocab]`` logits in the forward pass.

Replaces the reference's dependency on the cut-cross-entropy Triton kernels (``d9d/kernel/cce/cce.py:47-298``,
``main.py:119-225``) with epilogues of the hand-written tcgen05 GEMM:

* forward : logits tiles stay in TMEM; the epilogue reduces them to online-softmax partials and the target logit,
* backward: per token chunk, logits are recomputed and the epilogue writes ``g * (softmax - onehot)`` as bf16 into
  a reusable chunk buffer which feeds two more tcgen05 GEMMs (``dE = dL @ C``, ``dC += dL^T @ E``).
"""

from __future__ import annotations

import os

from typing import Any

import torch
from torch import nn
from torch.autograd import Function

from d9d_b200.core.autograd import GLOBAL_GRAD_CONTEXT, GradDirection

from .._native import grad_dtype_of, native_ops, on_gpu

IGNORE_INDEX = -100
# Backward materialises g*(softmax - onehot) for a chunk of tokens at a time (bf16).  180 GB of HBM make a large chunk
# affordable, and every extra chunk costs one more read-modify-write pass over the fp32 dC [V, K] accumulator plus the
# tail waves of three GEMMs, so the default covers 16k tokens x 150k vocabulary in one piece.
_CHUNK_BYTES = int(os.environ.get("D9D_CCE_CHUNK_BYTES", 6 << 30))


def linear_cross_entropy_reference(e, c, targets, bias=None, ignore_index=IGNORE_INDEX, softcap=None):
    logits = e.float() @ c.float().t()
    if bias is not None:
        logits = logits + bias.float()
    if softcap is not None:
        logits = torch.tanh(logits / softcap) * softcap
    lse = torch.logsumexp(logits, dim=-1)
    nll = torch.nn.functional.cross_entropy(logits, targets, ignore_index=ignore_index, reduction="none")
    return nll, lse


class _LinearCEFunction(Function):
    @staticmethod
    def forward(ctx: Any, e: torch.Tensor, c: torch.Tensor, targets: torch.Tensor, ignore_index: int):
        ops = native_ops()
        nll, lse = ops.ce_forward(e, c, targets, ignore_index)
        ctx.save_for_backward(e, c, targets, lse)
        ctx.ignore_index = ignore_index
        ctx.mark_non_differentiable(lse)
        return nll, lse

    @staticmethod
    def backward(ctx: Any, grad_nll: torch.Tensor, _grad_lse: torch.Tensor):  # type: ignore[override]
        e, c, targets, lse = ctx.saved_tensors
        ops = native_ops()
        T, K = e.shape
        V = c.shape[0]
        need_e = ctx.needs_input_grad[0] and GLOBAL_GRAD_CONTEXT.check_direction(GradDirection.inputs)
        need_c = ctx.needs_input_grad[1] and GLOBAL_GRAD_CONTEXT.check_direction(GradDirection.weight)
        g = grad_nll.float().contiguous()
        de = torch.empty_like(e) if need_e else None
        dc = torch.zeros(V, K, device=c.device, dtype=torch.float32) if need_c else None
        v_pitch = (V + 7) // 8 * 8  # TMA needs 16-byte aligned row pitch
        chunk = max(128, min(T, (_CHUNK_BYTES // (2 * v_pitch)) // 128 * 128))
        buf = torch.empty(chunk, v_pitch, device=e.device, dtype=torch.bfloat16)[:, :V]
        for t0 in range(0, T, chunk):
            t1 = min(t0 + chunk, T)
            dl = buf[: t1 - t0]
            ops.ce_dlogits(e[t0:t1], c, targets[t0:t1], lse[t0:t1], g[t0:t1], dl, ctx.ignore_index)
            if need_e:
                ops.gemm(dl, c, de[t0:t1], False, True, False)  # dL[t,V] @ C[V,K]
            if need_c:
                ops.gemm(dl, e[t0:t1], dc, True, True, True)  # dC += dL^T @ E
        if dc is not None:
            grad_dtype = grad_dtype_of(c)
            dc = dc if grad_dtype == torch.float32 else dc.to(grad_dtype)
        return de, dc, None, None


def linear_cross_entropy(
    e: torch.Tensor,
    c: torch.Tensor,
    targets: torch.Tensor,
    bias: torch.Tensor | None = None,
    ignore_index: int = IGNORE_INDEX,
    softcap: float | None = None,
    reduction: str = "mean",
    shift: bool | int = 0,
    return_lse: bool = False,
    **_unused: Any,
) -> torch.Tensor | tuple[torch.Tensor, torch.Tensor]:
    """``cross_entropy(e @ c^T, targets)`` with ``reduction in {none, mean, sum}``.

    ``e``: ``[..., K]`` embeddings, ``c``: ``[V, K]`` classifier, ``targets``: ``[...]`` int64.
    ``shift=n`` drops the last ``n`` embeddings / first ``n`` targets along the sequence dim (causal LM shift).
    Extra CCE-specific options of the reference (gradient filtering etc.) are accepted and ignored.
    """
    if isinstance(shift, bool):
        shift = int(shift)
    if shift:
        e = e[..., :-shift, :]
        targets = targets[..., shift:]
    lead_shape = targets.shape
    e2 = e.reshape(-1, e.shape[-1])
    t1 = targets.reshape(-1)
    native = on_gpu(e) and e.dtype == torch.bfloat16 and bias is None and softcap is None and e2.shape[-1] % 8 == 0
    if on_gpu(e) and not native and e.dtype == torch.bfloat16:
        # bias / softcap are rarely used by the models here; they take the (slow, materialising) oracle path explicitly
        native = False
    if native:
        nll, lse = _LinearCEFunction.apply(e2.contiguous(), c.to(e.dtype).contiguous(), t1.contiguous(), ignore_index)
    else:
        nll, lse = linear_cross_entropy_reference(e2, c, t1, bias, ignore_index, softcap)
    if reduction == "none":
        loss = nll.view(lead_shape)
    elif reduction == "sum":
        loss = nll.sum()
    elif reduction == "mean":
        loss = nll.sum() / (t1 != ignore_index).sum().clamp_min(1)
    else:
        raise ValueError(f"unknown reduction {reduction!r}")
    if return_lse:
        return loss, lse.view(lead_shape)
    return loss


class LinearCrossEntropy(nn.Module):
    """Module wrapper with stored options (reference ``d9d/kernel/cce/main.py:228-282``)."""

    def __init__(self, ignore_index: int = IGNORE_INDEX, softcap: float | None = None, reduction: str = "mean",
                 shift: bool | int = 0, **kwargs: Any):
        super().__init__()
        self.ignore_index, self.softcap, self.reduction, self.shift = ignore_index, softcap, reduction, shift
        self.kwargs = kwargs

    def forward(self, e: torch.Tensor, c: torch.Tensor, targets: torch.Tensor, bias: torch.Tensor | None = None):
        return linear_cross_entropy(e, c, targets, bias=bias, ignore_index=self.ignore_index, softcap=self.softcap,
                                    reduction=self.reduction, shift=self.shift)
