from __future__ import annotations

from collections.abc import Callable
from typing import Any

import torch
import torch.distributed as dist
from torch.autograd import Function

from d9d_b200.core.autograd import GLOBAL_GRAD_CONTEXT, GradDirection
from d9d_b200.internals.nvlink import SymmetricArena

from .._native import fused_wgrad_buffer, fused_wgrad_owner, grad_dtype_of, native_ops


# SMs the flag-gated GEMM leaves to the concurrent pull chain (measured: the small copy / flag / barrier kernels co-reside
# with the persistent GEMM CTAs, no SMs need to be set aside)
_PIPELINE_SPARE_SMS = 0
_DIRECT_PEER_LOAD_MAX_N = 512  # output width up to which A tiles are loaded directly from the peers (few re-reads)


class TensorParallelWorkspace:
    """Symmetric staging buffers of one tensor-parallel group (grown on demand; every rank grows in lock-step because
    all ranks execute the same layers with the same shapes)."""

    _instances: dict[str, "TensorParallelWorkspace"] = {}

    def __init__(self, group: dist.ProcessGroup):
        self.group = group
        self.world = group.size()
        self.rank = group.rank()
        self._arenas: dict[str, SymmetricArena] = {}
        self._side: torch.cuda.Stream | None = None
        self._warmed: set[tuple] = set()

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

    def gather_async(self, name: str, local: torch.Tensor, block_rows: int,
                     consumer: "Callable[[torch.Tensor, torch.Tensor], None] | None" = None) -> tuple[torch.Tensor, torch.Tensor]:
        """All-gather ``local [rows_local, C]`` into a *local* buffer ``[world * rows_local, C]`` (rows in
        ``(batch, rank, row-in-block)`` order) without blocking the current stream; returns ``(gathered, flags)``.

        The local shard is copied on the current stream, then ``consumer(gathered, flags)`` (typically the flag-gated
        GEMM) is launched, and only then the publish / pull chain is enqueued on a side stream: barrier → copy my shard
        to the symmetric staging buffer → barrier → pull every remote shard over NVLink exactly once → ``flags[r] = 1``.
        Enqueueing the consumer first matters when all streams share one hardware queue
        (``CUDA_DEVICE_MAX_CONNECTIONS=1``): work can only overlap with work submitted *earlier*.
        ``join()`` must be called before the buffer is used by anything that does not look at the flags.
        """
        rows_local, cols = local.shape
        arena = self.arena(name, local.numel(), local.device)
        batch = rows_local // block_rows
        gathered = torch.empty(self.world * rows_local, cols, device=local.device, dtype=local.dtype)
        view = gathered.view(batch, self.world, block_rows, cols)
        flags = torch.zeros(self.world, dtype=torch.int32, device=local.device)
        if self._side is None:
            self._side = torch.cuda.Stream(device=local.device)
        main = torch.cuda.current_stream()
        view[:, self.rank].copy_(local.view(batch, block_rows, cols))
        ready = torch.cuda.Event()
        ready.record(main)
        # CUDA loads kernels lazily and a first-time load waits for the device to drain: if the consumer is already
        # spinning on a flag that one of the not-yet-loaded side kernels must raise, that dead-locks.  The first call of
        # every kernel combination therefore runs the pull chain to completion before launching the consumer.
        warm_key = (name, batch > 1, local.dtype)
        warmed = warm_key in self._warmed
        if consumer is not None and warmed:
            consumer(gathered, flags)
        self._side.wait_event(ready)
        for t in (gathered, flags, local):
            t.record_stream(self._side)
        with torch.cuda.stream(self._side):
            arena.barrier()  # peers finished pulling the previous contents of my staging buffer
            arena.buffer[: local.numel()].copy_(local.reshape(-1))
            arena.barrier()  # every rank's shard is published
            for step in range(1, self.world):
                r = (self.rank + step) % self.world
                view[:, r].copy_(arena.peer_view(r, (rows_local * cols,))[: rows_local * cols].view(batch, block_rows, cols))
                flags[r : r + 1].fill_(1)
        if not warmed:
            self._warmed.add(warm_key)
            main.wait_stream(self._side)
            if consumer is not None:
                consumer(gathered, flags)
        return gathered, flags

    def join(self) -> None:
        if self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)

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
        y = torch.empty(rows_local * ws.world, weight.shape[0], device=x.device, dtype=x.dtype)
        if weight.shape[0] <= _DIRECT_PEER_LOAD_MAX_N:
            # few n-tiles: every A tile is read once or twice, load it straight from its owner through TMA
            _, ptrs = ws.stage("ag_in", x2)
            ops.gemm_ag_a(ptrs, rows_local, x2.shape[1], block_rows, weight, y, False)
        else:
            # many n-tiles re-read A: pull every remote shard once (copy engines) while the GEMM already runs on the
            # local shard; tiles of a remote shard wait for its arrival flag
            ws.gather_async("ag_in", x2, block_rows,
                            consumer=lambda g, f: ops.gemm_wait_a(g, f, ws.rank, block_rows, weight, y, False, _PIPELINE_SPARE_SMS))
            ws.join()
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
        need_dw = GLOBAL_GRAD_CONTEXT.check_direction(GradDirection.weight) and (ctx.needs_input_grad[1] or ctx.needs_input_grad[3])
        gathered = None
        if need_dw:  # re-gather x on the side stream; it overlaps the dgrad GEMM below
            gathered, _ = ws.gather_async("ag_in", x2, ctx.block_rows)
        if ctx.needs_input_grad[0] and GLOBAL_GRAD_CONTEXT.check_direction(GradDirection.inputs):
            # dx_shard = reduce_scatter(dy @ W): every rank reduce-adds its partial tiles into the owners' buffers
            arena, ptrs = ws.zeroed("rs_out", rows_local * k, dy.device)
            ops.gemm_rs_d(dy, weight, ptrs, rows_local, k, ctx.block_rows, True)
            arena.barrier()
            dx = arena.buffer[: rows_local * k].view(rows_local, k).clone().view(ctx.x_shape)
        if need_dw:  # dW[N, K] = dy^T @ all_gather(x)
            ws.join()
            if ctx.owner is not None and ctx.needs_input_grad[3]:
                ops.gemm(dy, gathered, fused_wgrad_buffer(ctx.owner), True, True, True)
            else:
                dw = torch.empty(weight.shape, device=weight.device, dtype=grad_dtype_of(weight))
                ops.gemm(dy, gathered, dw, True, True, False)
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
        if need_dx:  # dx[M, K_local] = all_gather(dy) @ W, consuming the shards as they land
            dx = torch.empty_like(x2)
            gathered, _ = ws.gather_async("ag_in", dy, ctx.block_rows,
                                          consumer=lambda g, f: ops.gemm_wait_a(g, f, ws.rank, ctx.block_rows, weight, dx, True, _PIPELINE_SPARE_SMS))
            dx = dx.view(ctx.x_shape)
        elif need_dw:
            gathered, _ = ws.gather_async("ag_in", dy, ctx.block_rows)
        if need_dx or need_dw:
            ws.join()
        if need_dw:  # dW[N, K_local] = all_gather(dy)^T @ x
            if ctx.owner is not None and ctx.needs_input_grad[3]:
                ops.gemm(gathered, x2, fused_wgrad_buffer(ctx.owner), True, True, True)
            else:
                dw = torch.empty(weight.shape, device=weight.device, dtype=grad_dtype_of(weight))
                ops.gemm(gathered, x2, dw, True, True, False)
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
