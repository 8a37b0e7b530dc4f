from __future__ import annotations

import dataclasses
from typing import Any

import torch
import torch.distributed as dist
from torch.autograd import Function

from d9d_b200.internals.nvlink import SymmetricArena
from d9d_b200.kernel._native import native_ops
from d9d_b200.kernel.moe import MoELayout, build_moe_layout
from d9d_b200.kernel.moe.layout import ALIGN, layout_capacity

from .base import ExpertCommunicationHandler


class _Workspace:
    """Symmetric staging memory of one expert-parallel group, shared by all MoE layers (phases are serialised by the
    device-side barriers): ``[rows x hidden bf16 | rows fp32 | world x experts int32]``."""

    _instances: dict[str, "_Workspace"] = {}

    def __init__(self, group: dist.ProcessGroup):
        self.group = group
        self.arena: SymmetricArena | None = None
        self.rows = self.hidden = self.num_experts = 0

    @classmethod
    def for_group(cls, group: dist.ProcessGroup) -> "_Workspace":
        ws = cls._instances.get(group.group_name)
        if ws is None:
            ws = cls._instances[group.group_name] = cls(group)
        return ws

    def ensure(self, rows: int, hidden: int, num_experts: int, device: torch.device) -> None:
        if self.arena is not None and rows <= self.rows and hidden == self.hidden and num_experts == self.num_experts:
            return
        self.rows, self.hidden, self.num_experts = max(rows, self.rows), hidden, num_experts
        self.off_x = 0
        self.off_p = self.rows * hidden * 2
        self.off_counts = self.off_p + self.rows * 4
        total_bytes = self.off_counts + self.group.size() * num_experts * 4
        self.arena = SymmetricArena((total_bytes + 1) // 2, torch.bfloat16, device, self.group)
        raw = self.arena.buffer.view(torch.uint8)
        self.x = raw[self.off_x : self.off_p].view(torch.bfloat16).view(self.rows, hidden)
        self.p = raw[self.off_p : self.off_counts].view(torch.float32)
        self.counts = raw[self.off_counts : self.off_counts + self.group.size() * num_experts * 4].view(torch.int32).view(self.group.size(), num_experts)

    def peer_counts(self, rank: int) -> torch.Tensor:
        raw = self.arena.peer_view(rank, (self.arena.buffer.numel(),)).view(torch.uint8)
        return raw[self.off_counts : self.off_counts + self.group.size() * self.num_experts * 4].view(torch.int32).view(self.group.size(), self.num_experts)


@dataclasses.dataclass
class _Plan:
    ws: _Workspace
    dest_rank: torch.Tensor  # [T*k] int32
    dest_row: torch.Tensor  # [T*k] int32
    layout: MoELayout  # receive-side layout of this rank
    num_tokens: int
    top_k: int
    hidden: int


class _Dispatch(Function):
    @staticmethod
    def forward(ctx: Any, x: torch.Tensor, probs: torch.Tensor, plan: _Plan):
        ws, ops = plan.ws, native_ops()
        ops.ep_push(x, probs.reshape(-1).float().contiguous(), plan.dest_rank, plan.dest_row, ws.arena.peer_ptrs_dev, ws.off_x, ws.off_p, plan.top_k)
        ws.arena.barrier()  # every peer's rows have landed
        cap = plan.layout.capacity
        xp, pp = ws.x[:cap].clone(), ws.p[:cap].clone()
        ctx.plan, ctx.probs_dtype = plan, probs.dtype
        return xp, pp

    @staticmethod
    def backward(ctx: Any, dxp: torch.Tensor, dpp: torch.Tensor):  # type: ignore[override]
        plan: _Plan = ctx.plan
        ws, ops = plan.ws, native_ops()
        cap = plan.layout.capacity
        ws.x[:cap].copy_(dxp)
        if dpp is None:
            ws.p[:cap].zero_()
        else:
            ws.p[:cap].copy_(dpp.float())
        ws.arena.barrier()  # every owner published the gradients of the rows it received
        dx, dprobs = ops.ep_pull_sum(ws.arena.peer_ptrs_dev, ws.off_x, ws.off_p, plan.dest_rank, plan.dest_row, plan.num_tokens, plan.top_k,
                                     plan.hidden, True)
        ws.arena.barrier()  # peers finished reading before the workspace is reused
        return dx, dprobs.to(ctx.probs_dtype), None


class _Combine(Function):
    @staticmethod
    def forward(ctx: Any, yp: torch.Tensor, plan: _Plan):
        ws, ops = plan.ws, native_ops()
        ws.x[: plan.layout.capacity].copy_(yp)
        ws.arena.barrier()  # all expert outputs are published
        y, _ = ops.ep_pull_sum(ws.arena.peer_ptrs_dev, ws.off_x, ws.off_p, plan.dest_rank, plan.dest_row, plan.num_tokens, plan.top_k,
                               plan.hidden, False)
        ws.arena.barrier()
        ctx.plan = plan
        return y

    @staticmethod
    def backward(ctx: Any, dy: torch.Tensor):  # type: ignore[override]
        plan: _Plan = ctx.plan
        ws, ops = plan.ws, native_ops()
        cap = plan.layout.capacity
        ws.x[:cap].zero_()  # pad rows must stay zero for the grouped wgrad
        ws.arena.barrier()
        ops.ep_push(dy.contiguous(), None, plan.dest_rank, plan.dest_row, ws.arena.peer_ptrs_dev, ws.off_x, ws.off_p, plan.top_k)
        ws.arena.barrier()
        return ws.x[:cap].clone(), None


class NvlinkExpertParallelCommunicationHandler(ExpertCommunicationHandler):
    """Expert-parallel dispatch / combine over NVLink peer memory — no all-to-all, no host synchronisation.

    1. the (token, slot) pairs are counted / stably sorted per *global* expert on the device;
    2. every rank writes its per-expert counts into all peers' workspaces (tiny peer-to-peer copies) → barrier;
    3. from the complete ``[source rank, expert]`` count matrix every rank derives, with a few device-side tensor ops,
       (a) the 128-row aligned, expert-sorted layout of the rows it is going to *receive* and (b) for each of its own
       pairs the owner rank and the exact row in the owner's layout;
    4. ``ep_push`` stores the rows (and routing probabilities) straight into the owners' GEMM-ready buffers;
    5. after the experts, ``ep_pull_sum`` lets every token gather and sum its ``top_k`` outputs from the owners.
    The backward passes are the mirrored kernels.  Worst-case sized buffers make every shape static.

    Plays the role of the reference's DeepEP handler (``moe/communications/deepep.py:57-222``).
    """

    def __init__(self, num_experts: int):
        self._num_experts = num_experts
        self._group: dist.ProcessGroup | None = None
        self._plan: _Plan | None = None

    def setup(self, group: dist.ProcessGroup, hidden_size: int, hidden_dtype: torch.dtype) -> None:
        if self._num_experts % group.size() != 0:
            raise ValueError(f"{self._num_experts} experts cannot be split across {group.size()} expert-parallel ranks")
        self._group = group

    def dispatch(self, hidden_states: torch.Tensor, topk_ids: torch.Tensor, topk_weights: torch.Tensor):
        group = self._group
        if group is None:
            raise ValueError("setup() was not called")
        world, rank = group.size(), group.rank()
        experts, local_experts = self._num_experts, self._num_experts // group.size()
        tokens, top_k = topk_ids.shape
        hidden = hidden_states.shape[-1]
        device = hidden_states.device
        capacity = layout_capacity(world * tokens, top_k, local_experts)
        ws = _Workspace.for_group(group)
        ws.ensure(capacity, hidden, experts, device)

        send = build_moe_layout(topk_ids, experts, align=1)  # counts / offsets / stable sorted position per pair
        ws.x[:capacity].zero_()  # pad rows of the aligned layout must be zero; precedes the peers' pushes (barrier below)
        ws.p[:capacity].zero_()
        for peer in range(world):
            ws.peer_counts(peer)[rank].copy_(send.counts)
        ws.arena.barrier()  # count matrix complete everywhere, workspaces cleared

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
        dest_row = torch.where(valid, base.reshape(-1)[safe] + within, torch.full_like(flat, -1)).int()
        dest_rank = (safe // local_experts).int()

        my_seg = seg[rank]
        tile_start = torch.arange(capacity // ALIGN, device=device) * ALIGN
        tile_owner = torch.searchsorted(my_seg, tile_start, right=True) - 1
        tile_group = torch.where(tile_start < my_seg[-1], tile_owner, torch.full_like(tile_owner, -1)).int()
        layout = MoELayout(counts=received[rank].int(), seg_offsets=my_seg.int(), row_map=torch.empty(0, dtype=torch.int32, device=device),
                           tile_group=tile_group, num_tokens=world * tokens, top_k=top_k, num_experts=local_experts, capacity=capacity)
        plan = _Plan(ws=ws, dest_rank=dest_rank, dest_row=dest_row, layout=layout, num_tokens=tokens, top_k=top_k, hidden=hidden)
        self._plan = plan
        xp, pp = _Dispatch.apply(hidden_states.contiguous(), topk_weights, plan)
        return xp, pp, layout

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

    def combine(self, hidden_states: torch.Tensor) -> torch.Tensor:
        if self._active is None:
            raise ValueError("Cannot run combine before running dispatch!")
        active, self._active = self._active, None
        return active.combine(hidden_states)
