"""The whole expert block ``probs * down(silu(gate(x)) * up(x))`` over the 128-row aligned layout as ONE autograd node.

Compared with composing ``grouped_linear`` / ``silu_mul_probs`` nodes this
* lets the two input-gradient GEMMs (gate, up) accumulate into one buffer in their epilogues (no separate add pass),
* limits the element-wise kernels to the rows that are really in use (static-capacity buffers of the expert-parallel path
  are mostly empty),
* lets the caller name the buffers the block writes its output / input gradient into - the expert-parallel handler passes
  its NVLink-mapped staging region, so the peers read the results in place (no staging copies),
* honours split backward passes (``GLOBAL_GRAD_CONTEXT``): the input pass stashes what the weight pass needs.

Reference: ``d9d/module/block/moe/grouped_experts.py:36-73`` (three ``gmm`` calls + ``silu_mul`` + a probability multiply).
"""

from __future__ import annotations

from typing import Any

import torch
from torch.autograd import Function

from d9d_b200.core.autograd import GLOBAL_GRAD_CONTEXT, GradDirection

from .._native import fused_wgrad_buffer, fused_wgrad_owner, grad_dtype_of, native_ops
from .layout import MoELayout


def _wgrad(ops, x: torch.Tensor, dy: torch.Tensor, weight: torch.Tensor, owner: torch.Tensor | None, layout: MoELayout,
           needs_weight: bool, needs_owner: bool) -> torch.Tensor | None:
    """dW[e] (+)= x_e^T dy_e; accumulates straight into the owner's pre-allocated gradient when there is one."""
    if owner is not None and needs_owner:
        ops.gemm_grouped_k(x, dy, fused_wgrad_buffer(owner), layout.seg_offsets, True)
        return None
    if needs_weight:
        dw = torch.empty(weight.shape, device=weight.device, dtype=grad_dtype_of(weight))
        ops.gemm_grouped_k(x, dy, dw, layout.seg_offsets, False)
        return dw
    return None


class _GroupedSwiGLUFunction(Function):
    @staticmethod
    def forward(ctx: Any, xp, pp, wg, wu, wd, layout: MoELayout, og, ou, od, out_buf, dx_buf):
        ops = native_ops()
        cap, ffn = xp.shape[0], wg.shape[2]
        valid = layout.seg_offsets[-1:]  # device scalar: rows in use
        gate = torch.empty(cap, ffn, device=xp.device, dtype=xp.dtype)
        up = torch.empty(cap, ffn, device=xp.device, dtype=xp.dtype)
        ops.gemm_grouped_m(xp, wg, gate, layout.tile_group, True)
        ops.gemm_grouped_m(xp, wu, up, layout.tile_group, True)
        h = ops.silu_mul_probs_fwd(gate, up, pp, valid)
        y = out_buf[:cap] if out_buf is not None else torch.empty(cap, wd.shape[2], device=xp.device, dtype=xp.dtype)
        ops.gemm_grouped_m(h, wd, y, layout.tile_group, True)
        ctx.save_for_backward(xp, pp, gate, up, h, wg, wu, wd)
        ctx.layout, ctx.owners, ctx.dx_buf, ctx.stash = layout, (og, ou, od), dx_buf, None
        return y

    @staticmethod
    def backward(ctx: Any, dy: torch.Tensor):  # type: ignore[override]
        xp, pp, gate, up, h, wg, wu, wd = ctx.saved_tensors
        layout: MoELayout = ctx.layout
        og, ou, od = ctx.owners
        ops = native_ops()
        do_inputs = GLOBAL_GRAD_CONTEXT.check_direction(GradDirection.inputs)
        do_weights = GLOBAL_GRAD_CONTEXT.check_direction(GradDirection.weight)
        valid = layout.seg_offsets[-1:]
        cap = xp.shape[0]

        if ctx.stash is not None:  # weight pass of a split backward: the input pass left what we need
            dy, dgate, dup = ctx.stash
            ctx.stash = None
        else:
            dy = dy.contiguous()
            dh = torch.empty(cap, wd.shape[1], device=xp.device, dtype=xp.dtype)
            ops.gemm_grouped_m(dy, wd, dh, layout.tile_group, False)  # W_down[e] read as [N' = ffn, K' = hidden]
            dgate = dup = None

        dwd = dwg = dwu = dxp = dpp = None
        if do_weights:
            dwd = _wgrad(ops, h, dy, wd, od, layout, ctx.needs_input_grad[4], ctx.needs_input_grad[8])
        if dgate is None:
            dgate, dup, dpp = ops.silu_mul_probs_bwd(dh, gate, up, pp, valid)
        if do_inputs and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]):
            if ctx.needs_input_grad[0]:
                whole = do_weights  # a split backward reads dy again later: never alias the caller's staging buffers then
                dxp = ctx.dx_buf[:cap] if (ctx.dx_buf is not None and whole) else torch.empty_like(xp)
                ops.gemm_grouped_m(dgate, wg, dxp, layout.tile_group, False)
                ops.gemm_grouped_m(dup, wu, dxp, layout.tile_group, False, True)  # += in the epilogue
        if do_weights:
            dwg = _wgrad(ops, xp, dgate, wg, og, layout, ctx.needs_input_grad[2], ctx.needs_input_grad[6])
            dwu = _wgrad(ops, xp, dup, wu, ou, layout, ctx.needs_input_grad[3], ctx.needs_input_grad[7])
        elif do_inputs:
            ctx.stash = (dy.clone(), dgate, dup)  # the weight pass follows later
        if not ctx.needs_input_grad[1]:
            dpp = None
        return dxp, dpp, dwg, dwu, dwd, None, None, None, None, None, None


def grouped_swiglu(xp: torch.Tensor, pp: torch.Tensor, w_gate: torch.Tensor, w_up: torch.Tensor, w_down: torch.Tensor,
                   layout: MoELayout, out: torch.Tensor | None = None, dx_out: torch.Tensor | None = None) -> torch.Tensor:
    """CUDA bf16 only.  ``w_*`` are ``[E, in, out]``; ``out`` / ``dx_out``: optional ``[>= capacity, hidden]`` buffers the
    block writes its result / its input gradient into (the returned tensor / the gradient then alias them)."""
    owners, weights = [], []
    for w in (w_gate, w_up, w_down):
        owner = fused_wgrad_owner(w)
        owners.append(owner)
        weights.append(w.detach() if owner is not None else w)
    return _GroupedSwiGLUFunction.apply(xp, pp.float(), *weights, layout, *owners, out, dx_out)
