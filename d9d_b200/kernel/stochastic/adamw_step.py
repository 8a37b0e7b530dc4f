from __future__ import annotations

import math
from collections.abc import Sequence

import torch

from .._native import native_ops, on_gpu
from ._rng import draw_seed, sr_round_reference

_CHUNK = 8192  # must match d9d::ADAM_CHUNK


def _validate(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor) -> None:
    if g.shape != p.shape:
        raise ValueError("Shape mismatch between grads and params.")
    if m.shape != p.shape:
        raise ValueError("Shape mismatch between exp_avg state and params.")
    if v.shape != p.shape:
        raise ValueError("Shape mismatch between exp_avg_sq state and params.")
    if p.dtype != torch.bfloat16:
        raise ValueError("Params must be BFloat16 for this kernel.")
    if not p.is_contiguous():
        raise ValueError("Params must be contiguous since it is an in-place kernel.")
    if not m.is_contiguous():
        raise ValueError("Exp_avg state must be contiguous since it is an in-place kernel.")
    if not v.is_contiguous():
        raise ValueError("Exp_avg_sq state must be contiguous since it is an in-place kernel.")
    if m.dtype != v.dtype:
        raise ValueError("States have different dtypes.")
    # the kernel distinguishes bf16 from "everything else = fp32": any other dtype would be reinterpreted silently
    if g.dtype not in (torch.bfloat16, torch.float32):
        raise ValueError(f"Grads must be BFloat16 or Float32 for this kernel, got {g.dtype}.")
    if m.dtype not in (torch.bfloat16, torch.float32):
        raise ValueError(f"Optimizer states must be BFloat16 or Float32 for this kernel, got {m.dtype}.")


class AdamWLaunchPlan:
    """Device-resident pointer table + block map for one multi-tensor launch (built once, reused every step)."""

    def __init__(self, params: Sequence[torch.Tensor], grads: Sequence[torch.Tensor], exp_avgs: Sequence[torch.Tensor],
                 exp_avg_sqs: Sequence[torch.Tensor]):
        metas, blocks, rng_base = [], [], 0
        for idx, (p, g, m, v) in enumerate(zip(params, grads, exp_avgs, exp_avg_sqs, strict=True)):
            n = p.numel()
            metas.append((p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, rng_base))
            blocks.extend((idx, c) for c in range((n + _CHUNK - 1) // _CHUNK))
            rng_base += (n + 7) // 8 * 8
        device = params[0].device
        self.key = tuple(m[:5] for m in metas)
        self.metas = torch.tensor(metas, dtype=torch.int64).reshape(-1, 6).to(device)
        self.block_map = torch.tensor(blocks, dtype=torch.int32).reshape(-1, 2).to(device)
        self.grad_bf16 = grads[0].dtype == torch.bfloat16
        self.state_bf16 = exp_avgs[0].dtype == torch.bfloat16
        # keep the tensors alive: the table stores raw pointers
        self._keepalive = (list(params), list(grads), list(exp_avgs), list(exp_avg_sqs))

    def matches(self, params, grads, exp_avgs, exp_avg_sqs) -> bool:
        if len(params) != len(self.key):
            return False
        return all(
            k == (p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel())
            for k, p, g, m, v in zip(self.key, params, grads, exp_avgs, exp_avg_sqs, strict=True)
        )


def _reference_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, seed, grad_scale):
    gf = g.float() * (1.0 if grad_scale is None else float(grad_scale))
    pf = p.float() * (1.0 - lr * weight_decay)
    mn = beta1 * m.float() + (1.0 - beta1) * gf
    vn = beta2 * v.float() + (1.0 - beta2) * gf * gf
    bc1 = 1.0 - math.exp(step * math.log(beta1)) if beta1 > 0 else 1.0
    bc2 = 1.0 - math.exp(step * math.log(beta2)) if beta2 > 0 else 1.0
    pf = pf - (lr * (mn / bc1)) / ((vn / bc2).sqrt() + eps)
    p.copy_(sr_round_reference(pf, seed))
    if m.dtype == torch.bfloat16:
        m.copy_(sr_round_reference(mn, seed + 42))
        v.copy_(sr_round_reference(vn, seed + 67))
    else:
        m.copy_(mn)
        v.copy_(vn)


def adamw_stochastic_bf16_multi_(
    params: Sequence[torch.Tensor],
    grads: Sequence[torch.Tensor],
    exp_avgs: Sequence[torch.Tensor],
    exp_avg_sqs: Sequence[torch.Tensor],
    *,
    lr: float,
    beta1: float,
    beta2: float,
    eps: float,
    weight_decay: float,
    step: int,
    seed: int | None = None,
    generator: torch.Generator | None = None,
    grad_scale: torch.Tensor | None = None,
    plan: AdamWLaunchPlan | None = None,
) -> AdamWLaunchPlan | None:
    """One fused launch updating every tensor in the lists (all grads share a dtype, all states share a dtype).

    ``grad_scale`` (device fp32 scalar) multiplies every gradient inside the kernel.  Returns the launch plan so
    callers can cache it (pointer tables are rebuilt only when a tensor moved).
    """
    if len(params) == 0:
        return plan
    for p, g, m, v in zip(params, grads, exp_avgs, exp_avg_sqs, strict=True):
        _validate(p, g, m, v)
    if seed is None:
        seed = draw_seed(generator)
    if not on_gpu(params[0]):
        for i, (p, g, m, v) in enumerate(zip(params, grads, exp_avgs, exp_avg_sqs, strict=True)):
            _reference_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, seed + 1000 * i, grad_scale)
        return None
    grads = [g if g.is_contiguous() else g.contiguous() for g in grads]
    if plan is None or not plan.matches(params, grads, exp_avgs, exp_avg_sqs):
        plan = AdamWLaunchPlan(params, grads, exp_avgs, exp_avg_sqs)
    native_ops().adamw_sr_multi_(plan.metas, plan.block_map, lr, beta1, beta2, eps, weight_decay, step, seed,
                                 grad_scale, plan.grad_bf16, plan.state_bf16)
    return plan


def adamw_stochastic_bf16_(
    params: torch.Tensor,
    grads: torch.Tensor,
    exp_avg: torch.Tensor,
    exp_avg_sq: torch.Tensor,
    lr: float,
    beta1: float,
    beta2: float,
    eps: float,
    weight_decay: float,
    step: int,
    generator: torch.Generator | None = None,
) -> None:
    """Single-tensor AdamW step with stochastic rounding (reference ``d9d/kernel/stochastic/adamw_step.py:97-197``)."""
    adamw_stochastic_bf16_multi_([params], [grads], [exp_avg], [exp_avg_sq], lr=lr, beta1=beta1, beta2=beta2, eps=eps,
                                 weight_decay=weight_decay, step=step, generator=generator)
