from __future__ import annotations

import dataclasses
import math
from typing import Any

import torch
import torch.distributed as dist
from torch.autograd import Function

from d9d_b200.internals.nvlink import SymmetricArena
from d9d_b200.kernel._native import native_ops
from d9d_b200.kernel.moe import MoELayout, build_moe_layout
from d9d_b200.kernel.moe.layout import ALIGN, layout_capacity

from .base import ExpertCommunicationHandler


def _capacity_factor() -> float:
    """``D9D_EP_CAPACITY_FACTOR``: receive capacity of a rank as a multiple of its fair share (``tokens * top_k`` rows).
    Unset / 0 = worst case (every pair of every rank lands here): always correct, but the static buffers are ``world`` times
    larger than what balanced routing needs.  With a factor, rows that do not fit are dropped and an overflow flag is raised
    (``NvlinkExpertParallelCommunicationHandler.overflowed``)."""
    import os

    try:
        return max(float(os.environ.get("D9D_EP_CAPACITY_FACTOR", "0")), 0.0)
    except ValueError:
        return 0.0


class _Region:
    """``[rows x hidden bf16 | rows fp32]`` in symmetric memory (same layout on every rank, mapped into every peer)."""

    def __init__(self, rows: int, hidden: int, group: dist.ProcessGroup, device: torch.device, extra_bytes: int = 0):
        self.rows, self.hidden = rows, hidden
        self.off_x = 0
        self.off_p = rows * hidden * 2
        self.off_extra = self.off_p + rows * 4
        total = self.off_extra + extra_bytes
        self.arena = SymmetricArena((total + 1) // 2, torch.bfloat16, device, group)
        raw = self.arena.buffer.view(torch.uint8)
        self.x = raw[self.off_x : self.off_p].view(torch.bfloat16).view(rows, hidden)
        self.p = raw[self.off_p : self.off_extra].view(torch.float32)
        self.raw = raw


class _Workspace:
    """Symmetric memory of one expert-parallel group shared by all MoE layers (phases are serialised by the device-side
    barriers): the per-expert count matrix plus one ``[rows x hidden | rows]`` staging region that carries expert outputs
    (combine forward), output gradients (combine backward) and input gradients (dispatch backward)."""

    _instances: dict[str, "_Workspace"] = {}

    def __init__(self, group: dist.ProcessGroup):
        self.group = group
        self.region: _Region | None = None
        self.num_experts = 0
        # receive buffers not in use by any layer.  Shared by all layers of the group: without activation checkpointing every
        # layer keeps its buffer from forward to backward (the pool just grows to one per layer and microbatch in flight),
        # with checkpointing a handful of buffers serve the whole model.
        self.pool: list[_Region] = []

    def acquire(self, rows: int, hidden: int, device: torch.device) -> "_Region":
        for i, r in enumerate(self.pool):
            if r.rows >= rows and r.hidden == hidden:
                return self.pool.pop(i)
        # collective allocation: every rank of the group runs the same schedule, so pools grow in lock-step
        return _Region(rows, hidden, self.group, device)

    @classmethod
    def for_group(cls, group: dist.ProcessGroup) -> "_Workspace":
        ws = cls._instances.get(group.group_name)
        if ws is None:
            ws = cls._instances[group.group_name] = cls(group)
        return ws

    def ensure(self, rows: int, hidden: int, num_experts: int, device: torch.device) -> None:
        r = self.region
        if r is not None and rows <= r.rows and hidden == r.hidden and num_experts == self.num_experts:
            return
        rows = max(rows, r.rows if r is not None else 0)
        self.num_experts = num_experts
        world = self.group.size()
        self.region = _Region(rows, hidden, self.group, device, extra_bytes=world * num_experts * 4)
        r = self.region
        self.counts = r.raw[r.off_extra : r.off_extra + world * num_experts * 4].view(torch.int32).view(world, num_experts)

    def peer_counts(self, rank: int) -> torch.Tensor:
        r = self.region
        world = self.group.size()
        raw = r.arena.peer_view(rank, (r.arena.buffer.numel(),)).view(torch.uint8)
        return raw[r.off_extra : r.off_extra + world * self.num_experts * 4].view(torch.int32).view(world, self.num_experts)

    def barrier(self) -> None:
        self.region.arena.barrier()


class _ReleaseOnDelete:
    """Runs ``release`` when the last reference (the autograd node's context) goes away."""

    def __init__(self, release: Any):
        self._release = release

    def __del__(self) -> None:
        try:
            self._release()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass


@dataclasses.dataclass
class _Plan:
    ws: _Workspace
    recv: _Region  # this layer's receive buffer: the owners' GEMM-ready rows live here until the backward pass is done
    release: Any  # callable returning ``recv`` to the handler's pool
    dest_rank: torch.Tensor  # [T*k] int32
    dest_row: torch.Tensor  # [T*k] int32 (-1: dropped)
    layout: MoELayout  # receive-side layout of this rank
    num_tokens: int
    top_k: int
    hidden: int
    guard: Any = None


class _Dispatch(Function):
    """Rows (and routing probabilities) are stored straight into the owners' per-layer receive buffers; the returned
    tensors ARE those buffers (no staging copy): they stay valid until this layer's backward has run."""

    @staticmethod
    def forward(ctx: Any, x: torch.Tensor, probs: torch.Tensor, plan: _Plan):
        ops, recv = native_ops(), plan.recv
        ops.ep_push(x, probs.reshape(-1).float().contiguous(), plan.dest_rank, plan.dest_row, recv.arena.peer_ptrs_dev, recv.off_x, recv.off_p,
                    plan.top_k)
        plan.ws.barrier()  # every peer's rows have landed
        cap = plan.layout.capacity
        ctx.plan, ctx.probs_dtype = plan, probs.dtype
        # a graph that is dropped without a backward pass (the recomputation of non-reentrant activation checkpointing,
        # an abandoned forward) must still hand its receive buffer back
        plan.guard = _ReleaseOnDelete(plan.release)
        return recv.x[:cap], recv.p[:cap]

    @staticmethod
    def backward(ctx: Any, dxp: torch.Tensor, dpp: torch.Tensor):  # type: ignore[override]
        plan: _Plan = ctx.plan
        ws, ops = plan.ws, native_ops()
        r, cap = ws.region, plan.layout.capacity
        if dxp.data_ptr() != r.x.data_ptr():  # the fused expert block writes its input gradient straight into the staging region
            r.x[:cap].copy_(dxp)
        if dpp is None:
            r.p[:cap].zero_()
        else:
            r.p[:cap].copy_(dpp.float())
        ws.barrier()  # every owner published the gradients of the rows it received
        dx, dprobs = ops.ep_pull_sum(r.arena.peer_ptrs_dev, r.off_x, r.off_p, plan.dest_rank, plan.dest_row, plan.num_tokens, plan.top_k,
                                     plan.hidden, True)
        # no trailing barrier: the next writer of the staging region (the combine backward of the previous layer, the next
        # forward) starts with one
        plan.release()
        return dx, dprobs.to(ctx.probs_dtype), None


class _Combine(Function):
    @staticmethod
    def forward(ctx: Any, yp: torch.Tensor, plan: _Plan):
        ws, ops = plan.ws, native_ops()
        r = ws.region
        if yp.data_ptr() != r.x.data_ptr():  # the fused expert block writes its output straight into the staging region
            r.x[: plan.layout.capacity].copy_(yp)
        ws.barrier()  # all expert outputs are published
        y, _ = ops.ep_pull_sum(r.arena.peer_ptrs_dev, r.off_x, r.off_p, plan.dest_rank, plan.dest_row, plan.num_tokens, plan.top_k,
                               plan.hidden, False)
        ctx.plan = plan
        if not torch.is_grad_enabled() or not yp.requires_grad:
            plan.release()
        return y

    @staticmethod
    def backward(ctx: Any, dy: torch.Tensor):  # type: ignore[override]
        plan: _Plan = ctx.plan
        ws, ops = plan.ws, native_ops()
        r, layout = ws.region, plan.layout
        ws.barrier()  # peers are done reading the staging region (previous phase)
        # pad rows of the aligned layout must be zero for the grouped wgrad; the peers only ever write real rows
        ops.moe_zero_pad(r.x, None, layout.counts, layout.seg_offsets)
        ops.ep_push(dy.contiguous(), None, plan.dest_rank, plan.dest_row, r.arena.peer_ptrs_dev, r.off_x, r.off_p, plan.top_k)
        ws.barrier()
        # a view of the staging region: consumed by the down-projection dgrad / wgrad before the dispatch backward of this
        # layer (stream order) overwrites it
        return r.x[: layout.capacity], None


class NvlinkExpertParallelCommunicationHandler(ExpertCommunicationHandler):
    """Expert-parallel dispatch / combine over NVLink peer memory — no all-to-all, no host synchronisation.

    1. the (token, slot) pairs are counted / stably sorted per *global* expert on the device;
    2. every rank writes its per-expert counts into all peers' workspaces (tiny peer-to-peer copies) → barrier;
    3. from the complete ``[source rank, expert]`` count matrix every rank derives, with a few device-side tensor ops,
       (a) the 128-row aligned, expert-sorted layout of the rows it is going to *receive* and (b) for each of its own
       pairs the owner rank and the exact row in the owner's layout;
    4. ``ep_push`` stores the rows (and routing probabilities) straight into the owners' GEMM-ready *per-layer* receive
       buffers - the grouped GEMMs read them in place and autograd keeps them as the saved activations (no staging copy);
    5. after the experts, ``ep_pull_sum`` lets every token gather and sum its ``top_k`` outputs from the owners.
    The backward passes are the mirrored kernels; 3 device barriers per layer and direction.  Buffers are static:
    worst case by default, ``D9D_EP_CAPACITY_FACTOR`` x the fair share otherwise (with an overflow flag).

    Plays the role of the reference's DeepEP handler (``moe/communications/deepep.py:57-222``).
    """

    instances: list["NvlinkExpertParallelCommunicationHandler"] = []  # for diagnostics (overflow checks)

    def __init__(self, num_experts: int):
        self._num_experts = num_experts
        self._group: dist.ProcessGroup | None = None
        self._plan: _Plan | None = None
        self._overflow: torch.Tensor | None = None
        NvlinkExpertParallelCommunicationHandler.instances.append(self)

    def setup(self, group: dist.ProcessGroup, hidden_size: int, hidden_dtype: torch.dtype) -> None:
        if self._num_experts % group.size() != 0:
            raise ValueError(f"{self._num_experts} experts cannot be split across {group.size()} expert-parallel ranks")
        self._group = group

    @property
    def overflowed(self) -> bool:
        """True if rows were dropped because a capacity factor was too small (host synchronisation)."""
        return bool(self._overflow.item()) if self._overflow is not None else False

    def dispatch(self, hidden_states: torch.Tensor, topk_ids: torch.Tensor, topk_weights: torch.Tensor):
        group = self._group
        if group is None:
            raise ValueError("setup() was not called")
        world, rank = group.size(), group.rank()
        experts, local_experts = self._num_experts, self._num_experts // group.size()
        tokens, top_k = topk_ids.shape
        hidden = hidden_states.shape[-1]
        device = hidden_states.device
        factor = _capacity_factor()
        capacity = layout_capacity(world * tokens, top_k, local_experts)
        if factor > 0:
            capacity = min(capacity, layout_capacity(int(math.ceil(tokens * factor)), top_k, local_experts))
        ws = _Workspace.for_group(group)
        ws.ensure(capacity, hidden, experts, device)
        recv = ws.acquire(capacity, hidden, device)

        send = build_moe_layout(topk_ids, experts, align=1)  # counts / offsets / stable sorted position per pair
        for peer in range(world):
            ws.peer_counts(peer)[rank].copy_(send.counts)
        ws.barrier()  # count matrix complete everywhere; every reader of the staging region has passed its pull

        per_dest = ws.counts.view(world, world, local_experts).long()  # [source, owner, local expert]
        received = per_dest.sum(0)  # [owner, local expert]
        aligned = (received + ALIGN - 1) // ALIGN * ALIGN
        seg = torch.zeros(world, local_experts + 1, dtype=torch.long, device=device)
        seg[:, 1:] = aligned.cumsum(1)
        before_me = per_dest[:rank].sum(0)  # rows of lower-ranked sources come first inside every expert segment
        base = seg[:, :-1] + before_me  # [owner, local expert] first row of *my* rows

        flat = topk_ids.reshape(-1).long()
        valid = (flat >= 0) & (flat < experts)
        safe = flat.clamp(0, experts - 1)
        within = send.row_map.long() - send.seg_offsets.long()[safe]  # position among my pairs of that expert
        dest_row = base.reshape(-1)[safe] + within
        if factor > 0:
            valid &= dest_row < capacity  # rows that do not fit the owner's static buffer are dropped ...
            over = (seg[:, -1] > capacity).any().reshape(1)  # ... and reported
            self._overflow = over if self._overflow is None else (self._overflow | over)
        dest_row = torch.where(valid, dest_row, torch.full_like(flat, -1)).int()
        dest_rank = (safe // local_experts).int()

        my_seg = seg[rank].clamp(max=capacity) if factor > 0 else seg[rank]
        my_counts = received[rank]
        if factor > 0:
            my_counts = torch.minimum(my_counts, (my_seg[1:] - my_seg[:-1]))
        tile_start = torch.arange(capacity // ALIGN, device=device) * ALIGN
        tile_owner = torch.searchsorted(my_seg, tile_start, right=True) - 1
        tile_group = torch.where(tile_start < my_seg[-1], tile_owner, torch.full_like(tile_owner, -1)).int()
        layout = MoELayout(counts=my_counts.int(), seg_offsets=my_seg.int(), row_map=torch.empty(0, dtype=torch.int32, device=device),
                           tile_group=tile_group, num_tokens=world * tokens, top_k=top_k, num_experts=local_experts, capacity=capacity)
        # pad rows of the aligned layout must be zero; the peers' pushes only ever write real rows, so no ordering is needed
        native_ops().moe_zero_pad(recv.x, recv.p, layout.counts, layout.seg_offsets)

        released = []

        def release() -> None:
            if not released:
                released.append(True)
                ws.pool.append(recv)

        plan = _Plan(ws=ws, recv=recv, release=release, dest_rank=dest_rank, dest_row=dest_row, layout=layout, num_tokens=tokens,
                     top_k=top_k, hidden=hidden)
        self._plan = plan
        xp, pp = _Dispatch.apply(hidden_states.contiguous(), topk_weights, plan)
        return xp, pp, layout

    def expert_buffers(self) -> tuple[torch.Tensor | None, torch.Tensor | None]:
        if self._plan is None:
            return None, None
        staging = self._plan.ws.region.x
        return staging, staging

    def combine(self, hidden_states: torch.Tensor) -> torch.Tensor:
        if self._plan is None:
            raise ValueError("Cannot run combine before running dispatch!")
        plan, self._plan = self._plan, None
        return _Combine.apply(hidden_states, plan)


class AutoExpertParallelCommunicationHandler(ExpertCommunicationHandler):
    """Chooses per call: NVLink peer-memory exchange for bf16 CUDA activations inside one NVLink domain (<= 8 ranks),
    the NCCL / gloo all-to-all handler otherwise (CPU tests, other dtypes, ``D9D_EP_NVLINK=0``).  The choice cannot be
    made when the layer is parallelised because parameters may still live on the meta device then."""

    def __init__(self, num_experts: int):
        from .expert_parallel import ExpertParallelCommunicationHandler

        self._nvlink = NvlinkExpertParallelCommunicationHandler(num_experts)
        self._collective = ExpertParallelCommunicationHandler(num_experts)
        self._active: ExpertCommunicationHandler | None = None
        self._world = 1

    def setup(self, group: dist.ProcessGroup, hidden_size: int, hidden_dtype: torch.dtype) -> None:
        self._world = group.size()
        self._nvlink.setup(group, hidden_size, hidden_dtype)
        self._collective.setup(group, hidden_size, hidden_dtype)

    def dispatch(self, hidden_states: torch.Tensor, topk_ids: torch.Tensor, topk_weights: torch.Tensor):
        import os

        fast = (hidden_states.is_cuda and hidden_states.dtype == torch.bfloat16 and hidden_states.shape[-1] % 8 == 0
                and self._world <= 8 and os.environ.get("D9D_EP_NVLINK", "1") != "0")
        self._active = self._nvlink if fast else self._collective
        return self._active.dispatch(hidden_states, topk_ids, topk_weights)

    def expert_buffers(self) -> tuple[torch.Tensor | None, torch.Tensor | None]:
        return self._active.expert_buffers() if self._active is not None else (None, None)

    def combine(self, hidden_states: torch.Tensor) -> torch.Tensor:
        if self._active is None:
            raise ValueError("Cannot run combine before running dispatch!")
        active, self._active = self._active, None
        return active.combine(hidden_states)
