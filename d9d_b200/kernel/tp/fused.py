from __future__ import annotations

from typing import Any

import torch
import torch.distributed as dist
from torch.autograd import Function

from d9d_b200.core.autograd import GLOBAL_GRAD_CONTEXT, GradDirection
from d9d_b200.internals.nvlink import SymmetricArena

from .._native import fused_wgrad_buffer, fused_wgrad_owner, grad_dtype_of, native_ops


class TensorParallelWorkspace:
    """Symmetric staging buffers of one tensor-parallel group (grown on demand; every rank grows in lock-step because
    all ranks execute the same layers with the same shapes)."""

    _instances: dict[str, "TensorParallelWorkspace"] = {}

    def __init__(self, group: dist.ProcessGroup):
        self.group = group
        self.world = group.size()
        self.rank = group.rank()
        self._arenas: dict[str, SymmetricArena] = {}

    @classmethod
    def for_group(cls, group: dist.ProcessGroup) -> "TensorParallelWorkspace":
        ws = cls._instances.get(group.group_name)
        if ws is None:
            ws = cls(group)
            cls._instances[group.group_name] = ws
        return ws

    def arena(self, name: str, numel: int, device: torch.device) -> SymmetricArena:
        cur = self._arenas.get(name)
        if cur is None or cur.buffer.numel() < numel:
            cur = SymmetricArena(max(numel, 1 << 20), torch.bfloat16, device, self.group)
            self._arenas[name] = cur
        return cur

    def stage(self, name: str, tensor: torch.Tensor) -> tuple[SymmetricArena, list[int]]:
        """Publish ``tensor`` (bf16, any shape) to the peers: barrier (peers are done with the previous contents) →
        copy into this rank's symmetric buffer → barrier (every rank's copy is visible)."""
        arena = self.arena(name, tensor.numel(), tensor.device)
        arena.barrier()
        arena.buffer[: tensor.numel()].copy_(tensor.reshape(-1))
        arena.barrier()
        return arena, [int(p) for p in arena.handle.buffer_ptrs]

    def zeroed(self, name: str, numel: int, device: torch.device) -> tuple[SymmetricArena, list[int]]:
        arena = self.arena(name, numel, device)
        arena.barrier()
        arena.buffer[:numel].zero_()
        arena.barrier()
        return arena, [int(p) for p in arena.handle.buffer_ptrs]


def _rows(x: torch.Tensor) -> tuple[torch.Tensor, int, int]:
    """``[..., S_local, C]`` -> (2-D view, rows per sequence block, number of rows)."""
    x2 = x.reshape(-1, x.shape[-1])
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    return x2, x.shape[-2], x2.shape[0]


class _AllGatherLinear(Function):
    """``y[..., S, N] = all_gather(x[..., S/W, K], dim=-2) @ weight[N, K]^T`` with the gather fused into the GEMM."""

    @staticmethod
    def forward(ctx: Any, x: torch.Tensor, weight: torch.Tensor, ws: TensorParallelWorkspace, owner: torch.Tensor | None):
        ops = native_ops()
        x2, block_rows, rows_local = _rows(x)
        _, ptrs = ws.stage("ag_in", x2)
        y = torch.empty(rows_local * ws.world, weight.shape[0], device=x.device, dtype=x.dtype)
        ops.gemm_ag_a(ptrs, rows_local, x2.shape[1], block_rows, weight, y, False)
        ctx.save_for_backward(x2, weight)
        ctx.ws, ctx.owner, ctx.block_rows, ctx.x_shape = ws, owner, block_rows, x.shape
        return y.view(*x.shape[:-2], block_rows * ws.world, weight.shape[0])

    @staticmethod
    def backward(ctx: Any, grad_out: torch.Tensor):  # type: ignore[override]
        ops = native_ops()
        x2, weight = ctx.saved_tensors
        ws: TensorParallelWorkspace = ctx.ws
        dy = grad_out.reshape(-1, grad_out.shape[-1])
        if not dy.is_contiguous():
            dy = dy.contiguous()
        rows_local, k = x2.shape
        dx = dw = None
        if ctx.needs_input_grad[0] and GLOBAL_GRAD_CONTEXT.check_direction(GradDirection.inputs):
            # dx_shard = reduce_scatter(dy @ W): every rank reduce-adds its partial tiles into the owners' buffers
            arena, ptrs = ws.zeroed("rs_out", rows_local * k, dy.device)
            ops.gemm_rs_d(dy, weight, ptrs, rows_local, k, ctx.block_rows, True)
            arena.barrier()
            dx = arena.buffer[: rows_local * k].view(rows_local, k).clone().view(ctx.x_shape)
        if GLOBAL_GRAD_CONTEXT.check_direction(GradDirection.weight) and (ctx.needs_input_grad[1] or ctx.needs_input_grad[3]):
            _, xptrs = ws.stage("ag_in", x2)  # dW[N, K] = dy^T @ all_gather(x): x tiles come from their owners
            if ctx.owner is not None and ctx.needs_input_grad[3]:
                ops.gemm_ag_k(dy, xptrs, False, rows_local, k, ctx.block_rows, fused_wgrad_buffer(ctx.owner), True)
            else:
                dw = torch.empty(weight.shape, device=weight.device, dtype=grad_dtype_of(weight))
                ops.gemm_ag_k(dy, xptrs, False, rows_local, k, ctx.block_rows, dw, False)
        return dx, dw, None, None


class _LinearReduceScatter(Function):
    """``y[..., S/W, N] = reduce_scatter(x[..., S, K_local] @ weight[N, K_local]^T, dim=-2)`` fused into the GEMM."""

    @staticmethod
    def forward(ctx: Any, x: torch.Tensor, weight: torch.Tensor, ws: TensorParallelWorkspace, owner: torch.Tensor | None):
        ops = native_ops()
        x2, seq, rows = _rows(x)
        block_rows = seq // ws.world
        rows_local = rows // ws.world
        n = weight.shape[0]
        arena, ptrs = ws.zeroed("rs_out", rows_local * n, x.device)
        ops.gemm_rs_d(x2, weight, ptrs, rows_local, n, block_rows, False)
        arena.barrier()
        y = arena.buffer[: rows_local * n].view(rows_local, n).clone()
        ctx.save_for_backward(x2, weight)
        ctx.ws, ctx.owner, ctx.block_rows, ctx.x_shape = ws, owner, block_rows, x.shape
        return y.view(*x.shape[:-2], block_rows, n)

    @staticmethod
    def backward(ctx: Any, grad_out: torch.Tensor):  # type: ignore[override]
        ops = native_ops()
        x2, weight = ctx.saved_tensors
        ws: TensorParallelWorkspace = ctx.ws
        dy = grad_out.reshape(-1, grad_out.shape[-1])
        if not dy.is_contiguous():
            dy = dy.contiguous()
        rows_local, n = dy.shape
        dx = dw = None
        need_dx = ctx.needs_input_grad[0] and GLOBAL_GRAD_CONTEXT.check_direction(GradDirection.inputs)
        need_dw = GLOBAL_GRAD_CONTEXT.check_direction(GradDirection.weight) and (ctx.needs_input_grad[1] or ctx.needs_input_grad[3])
        if need_dx or need_dw:
            _, ptrs = ws.stage("ag_in", dy)
        if need_dx:  # dx[M, K_local] = all_gather(dy) @ W
            dx = torch.empty_like(x2)
            ops.gemm_ag_a(ptrs, rows_local, n, ctx.block_rows, weight, dx, True)
            dx = dx.view(ctx.x_shape)
        if need_dw:  # dW[N, K_local] = all_gather(dy)^T @ x
            if ctx.owner is not None and ctx.needs_input_grad[3]:
                ops.gemm_ag_k(x2, ptrs, True, rows_local, n, ctx.block_rows, fused_wgrad_buffer(ctx.owner), True)
            else:
                dw = torch.empty(weight.shape, device=weight.device, dtype=grad_dtype_of(weight))
                ops.gemm_ag_k(x2, ptrs, True, rows_local, n, ctx.block_rows, dw, False)
        return dx, dw, None, None


def _apply(fn: type[Function], x: torch.Tensor, weight: torch.Tensor, group: dist.ProcessGroup) -> torch.Tensor:
    ws = TensorParallelWorkspace.for_group(group)
    owner = fused_wgrad_owner(weight)
    if owner is None:
        return fn.apply(x, weight, ws, None)
    return fn.apply(x, weight.detach(), ws, owner)


def fused_tp_supported(x: torch.Tensor, weight: torch.Tensor, group: dist.ProcessGroup, gathered_seq: int) -> bool:
    """Shapes the fused kernels accept: bf16 CUDA tensors, per-rank sequence blocks that are multiples of one M tile."""
    world = group.size()
    return (x.is_cuda and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and x.dim() >= 2 and world <= 8
            and gathered_seq % world == 0 and (gathered_seq // world) % 128 == 0 and x.shape[-1] % 8 == 0 and weight.shape[0] % 8 == 0)


def all_gather_linear(x: torch.Tensor, weight: torch.Tensor, group: dist.ProcessGroup) -> torch.Tensor:
    return _apply(_AllGatherLinear, x, weight, group)


def linear_reduce_scatter(x: torch.Tensor, weight: torch.Tensor, group: dist.ProcessGroup) -> torch.Tensor:
    return _apply(_LinearReduceScatter, x, weight, group)
