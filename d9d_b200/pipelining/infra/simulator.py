"""Dependency-driven list scheduler that *derives* pipeline programs instead of hard-coding phase formulas.

The reference writes every schedule as a hand-derived sequence of phases (``program/interleaved.py``,
``zerobubblev.py``, ``dualpipev.py``).  Here a schedule is a *policy* (what to prefer when several actions are
ready, how many forwards may be in flight) fed to one simulator that advances all ranks tick by tick and only ever
emits actions whose data dependencies are satisfied:

    F(s, m)      needs F(s-1, m)
    B/I(s, m)    needs F(s, m) and B/I(s+1, m)         (the last stage only needs its own forward)
    W(s, m)      needs I(s, m)

Results produced at tick t are visible to other actions from tick t+1 (one tick of transfer latency), which is
also what makes the later communication injection deadlock-free.  Any policy yields a *valid* program by
construction; policies only shape bubbles and memory.
"""

from __future__ import annotations

import dataclasses
from collections.abc import Callable

from .action import Action, ActionKind, AnyAction, ComposeAction, Program


@dataclasses.dataclass
class SchedulePolicy:
    split_backward: bool  # emit I + W instead of B
    #: max number of microbatches a rank may have *injected* into its first local stage whose backward(-input) at
    #: that stage is still outstanding; forwards of deeper local stages are never throttled (they drain the pipe
    #: towards the backward, so throttling them could deadlock).  None = unbounded
    max_inflight: Callable[[int], int | None]
    #: ordering key among ready forwards of a rank: lower runs first
    forward_key: Callable[[int, int], tuple]
    #: ordering key among ready backwards
    backward_key: Callable[[int, int], tuple]
    prefer_backward: bool = True  # 1F1B flavour: drain backwards before starting new forwards
    pair_forward_backward: bool = False  # DualPipeV: issue a ready F together with a ready B/I in one slot
    forward_only: bool = False
    eager_weight: bool = False  # run W right after its I when nothing better is ready (default: fill bubbles only)


def simulate(num_stages: int, num_microbatches: int, stage_to_rank: dict[int, int], policy: SchedulePolicy) -> tuple[Program, list[list[AnyAction | None]]]:
    """Returns ``(program per rank, timeline[tick][rank])`` of compute actions."""
    ranks = sorted(set(stage_to_rank.values()))
    stages_of = {r: [s for s in range(num_stages) if stage_to_rank[s] == r] for r in ranks}
    last = num_stages - 1
    M = num_microbatches

    done_f: dict[tuple[int, int], int] = {}  # (stage, mb) -> tick finished
    done_i: dict[tuple[int, int], int] = {}
    done_w: dict[tuple[int, int], int] = {}
    program: Program = {r: [] for r in ranks}
    timeline: list[list[AnyAction | None]] = []

    total_f = num_stages * M
    total_bwd = 0 if policy.forward_only else num_stages * M
    tick = 0
    guard = 0
    while len(done_f) < total_f or len(done_i) < total_bwd or (policy.split_backward and len(done_w) < total_bwd):
        row: list[AnyAction | None] = []
        staged: list[tuple[dict, tuple[int, int]]] = []
        for r in ranks:
            def visible(table: dict[tuple[int, int], int], key: tuple[int, int]) -> bool:
                return key in table and table[key] < tick

            ready_f = [
                (s, m) for s in stages_of[r] for m in range(M)
                if (s, m) not in done_f and (s == 0 or visible(done_f, (s - 1, m)))
            ]
            ready_b = [] if policy.forward_only else [
                (s, m) for s in stages_of[r] for m in range(M)
                if (s, m) not in done_i and visible(done_f, (s, m)) and (s == last or visible(done_i, (s + 1, m)))
            ]
            ready_w = [
                (s, m) for s in stages_of[r] for m in range(M)
                if policy.split_backward and (s, m) not in done_w and visible(done_i, (s, m))
            ]
            entry = stages_of[r][0]
            inflight = sum(1 for m in range(M) if (entry, m) in done_f and (entry, m) not in done_i)
            limit = policy.max_inflight(r)
            if not (policy.forward_only or limit is None or inflight < limit):
                ready_f = [sm for sm in ready_f if sm[0] != entry]
            can_forward = bool(ready_f)
            ready_f.sort(key=lambda sm: policy.forward_key(*sm))
            ready_b.sort(key=lambda sm: policy.backward_key(*sm))
            ready_w.sort(key=lambda sm: (sm[1], -sm[0]))

            bkind = ActionKind.BACKWARD_INPUT if policy.split_backward else ActionKind.BACKWARD_FULL
            chosen: AnyAction | None = None
            if policy.pair_forward_backward and can_forward and ready_b:
                f, b = ready_f[0], ready_b[0]
                chosen = ComposeAction((Action(ActionKind.FORWARD, *f), Action(bkind, *b)))
                staged += [(done_f, f), (done_i, b)]
            elif ready_b and (policy.prefer_backward or not can_forward):
                b = ready_b[0]
                chosen = Action(bkind, *b)
                staged.append((done_i, b))
            elif can_forward:
                f = ready_f[0]
                chosen = Action(ActionKind.FORWARD, *f)
                staged.append((done_f, f))
            elif ready_b:
                b = ready_b[0]
                chosen = Action(bkind, *b)
                staged.append((done_i, b))
            elif ready_w:
                w = ready_w[0]
                chosen = Action(ActionKind.BACKWARD_WEIGHT, *w)
                staged.append((done_w, w))
            if chosen is None and ready_w:
                w = ready_w[0]
                chosen = Action(ActionKind.BACKWARD_WEIGHT, *w)
                staged.append((done_w, w))
            if chosen is not None:
                program[r].append(chosen)
            row.append(chosen)
        for table, key in staged:
            table[key] = tick
        timeline.append(row)
        tick += 1
        guard = guard + 1 if all(a is None for a in row) else 0
        if guard > 2:
            raise RuntimeError("Deadlock while deriving the pipeline schedule (policy admits no ready action)")
    return program, timeline
