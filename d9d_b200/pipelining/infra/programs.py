"""Program builders of the six supported schedules (reference ``infra/schedule/program/*.py``).

``inference`` / ``gpipe`` / ``looped_bfs`` are written out directly (they *are* their definition: all forwards
stage by stage, then all backwards in reverse); the 1F1B family, ZBV and DualPipeV are policies of the generic
simulator in :mod:`.simulator`.
"""

from __future__ import annotations

import abc

from .action import Action, ActionKind, AnyAction, Program
from .communications import add_communication_ops, validate_program
from .simulator import SchedulePolicy, simulate
from .topology import ScheduleStyle, build_stage_to_host_rank_topology


class PipelineProgramBuilder(abc.ABC):
    """Builds the per-rank action lists of a schedule."""

    @property
    @abc.abstractmethod
    def num_stages_per_rank(self) -> int: ...

    @property
    @abc.abstractmethod
    def topology_style(self) -> ScheduleStyle: ...

    @property
    def has_backward(self) -> bool:
        return True

    @abc.abstractmethod
    def _timeline(self, num_microbatches: int, pp_size: int, stage_to_rank: dict[int, int]) -> list[list[AnyAction | None]]: ...

    def compose(self, num_microbatches: int, pp_size: int) -> Program:
        num_stages = self.num_stages_per_rank * pp_size
        stage_to_rank = build_stage_to_host_rank_topology(num_stages, pp_size, self.topology_style)
        timeline = self._timeline(num_microbatches, pp_size, stage_to_rank)
        program = add_communication_ops(timeline, list(range(pp_size)), stage_to_rank, num_stages)
        validate_program(program, stage_to_rank, num_stages, num_microbatches, self.has_backward)
        return program


class LoopedBFSPipelineProgramBuilder(PipelineProgramBuilder):
    """Breadth-first: for every local stage all microbatch forwards, then (training) all backwards in reverse stage
    and reverse microbatch order.  ``num_stages_per_rank=1`` is GPipe; ``inference_mode`` drops the backward."""

    def __init__(self, num_stages_per_rank: int, inference_mode: bool = False):
        self._spr = num_stages_per_rank
        self._inference = inference_mode

    @property
    def num_stages_per_rank(self) -> int:
        return self._spr

    @property
    def topology_style(self) -> ScheduleStyle:
        return ScheduleStyle.loop

    @property
    def has_backward(self) -> bool:
        return not self._inference

    def _timeline(self, num_microbatches, pp_size, stage_to_rank):
        num_stages = self._spr * pp_size
        policy = SchedulePolicy(
            split_backward=False,
            max_inflight=lambda r: None,
            forward_key=lambda s, m: (s, m),
            backward_key=lambda s, m: (-s, -m),
            prefer_backward=False,  # all forwards first
            forward_only=self._inference,
        )
        _, timeline = simulate(num_stages, num_microbatches, stage_to_rank, policy)
        return timeline


class Interleaved1F1BPipelineProgramBuilder(PipelineProgramBuilder):
    """(Interleaved) one-forward-one-backward; with ``enable_zero_bubble`` the backward is split into I and W and the
    deferred W fill the bubbles (ZB1P flavour)."""

    def __init__(self, num_stages_per_rank: int, enable_zero_bubble: bool = False):
        self._spr = num_stages_per_rank
        self._zb = enable_zero_bubble

    @property
    def num_stages_per_rank(self) -> int:
        return self._spr

    @property
    def topology_style(self) -> ScheduleStyle:
        return ScheduleStyle.loop

    def _timeline(self, num_microbatches, pp_size, stage_to_rank):
        P, V, M = pp_size, self._spr, num_microbatches
        num_stages = P * V
        group = max(1, min(P, M))  # microbatches advance through the virtual stages in groups of P

        def limit(rank: int) -> int:
            # microbatches a rank may have in flight at its first chunk: the classic P - rank for plain 1F1B; with
            # interleaving a whole group of P must be admitted before the second chunk can start
            return (P - rank) if V == 1 else (2 * P - 1 - rank)

        policy = SchedulePolicy(
            split_backward=self._zb,
            max_inflight=limit,
            forward_key=lambda s, m: (m // group, s // P, m % group),
            backward_key=lambda s, m: (m // group, -(s // P), m % group),
            prefer_backward=True,
        )
        _, timeline = simulate(num_stages, M, stage_to_rank, policy)
        return timeline


class ZeroBubbleVPipelineProgramBuilder(PipelineProgramBuilder):
    """Zero-bubble schedule on the V topology (two stages per rank, first and last stage share rank 0) with split
    backward; memory is bounded like 1F1B (at most ~2P forwards in flight per rank)."""

    @property
    def num_stages_per_rank(self) -> int:
        return 2

    @property
    def topology_style(self) -> ScheduleStyle:
        return ScheduleStyle.v

    def _timeline(self, num_microbatches, pp_size, stage_to_rank):
        P, M = pp_size, num_microbatches
        policy = SchedulePolicy(
            split_backward=True,
            max_inflight=lambda r: 2 * P - r,
            forward_key=lambda s, m: (m, s),
            backward_key=lambda s, m: (m, -s),
            prefer_backward=True,
        )
        _, timeline = simulate(2 * P, M, stage_to_rank, policy)
        return timeline


class DualPipeVPipelineProgramBuilder(PipelineProgramBuilder):
    """DualPipeV: V topology, split backward, and forward/backward of different microbatches paired in one slot in
    the steady state (so an implementation can overlap one's compute with the other's communication).
    Requires ``num_microbatches >= 2 * pp_size`` like the reference."""

    @property
    def num_stages_per_rank(self) -> int:
        return 2

    @property
    def topology_style(self) -> ScheduleStyle:
        return ScheduleStyle.v

    def _timeline(self, num_microbatches, pp_size, stage_to_rank):
        P, M = pp_size, num_microbatches
        if M < 2 * P:
            raise ValueError(f"DualPipeV needs num_microbatches >= 2 * pp_size ({M} < {2 * P})")
        policy = SchedulePolicy(
            split_backward=True,
            max_inflight=lambda r: 2 * P - r + 1,
            forward_key=lambda s, m: (m, s),
            backward_key=lambda s, m: (m, -s),
            prefer_backward=True,
            pair_forward_backward=True,
        )
        _, timeline = simulate(2 * P, M, stage_to_rank, policy)
        return timeline


def program_to_text(program: Program) -> str:
    """Human-readable dump (one line per rank)."""
    return "\n".join(f"rank {r}: " + " ".join(str(a) for a in actions) for r, actions in sorted(program.items()))


__all__ = [
    "Action",
    "ActionKind",
    "DualPipeVPipelineProgramBuilder",
    "Interleaved1F1BPipelineProgramBuilder",
    "LoopedBFSPipelineProgramBuilder",
    "PipelineProgramBuilder",
    "ZeroBubbleVPipelineProgramBuilder",
    "program_to_text",
]
