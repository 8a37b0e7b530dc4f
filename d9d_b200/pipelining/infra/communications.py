"""Injection of SEND/RECV actions into a compute-only program.

Walk the compute timeline in global order; right after a compute action whose result is consumed on another rank
emit the SEND on the producer and — at that very moment — append the matching RECV to the consumer's list.  Every
rank's communication sequence is therefore a subsequence of ONE global order of (send, recv) pairs, which rules out
cyclic waits between blocking point-to-point operations (same guarantee as the reference's round-robin simulation,
``component/program/communications.py:119-189``).
"""

from __future__ import annotations

from .action import Action, ActionKind, AnyAction, Program, flatten


def add_communication_ops(timeline: list[list[AnyAction | None]], ranks: list[int], stage_to_rank: dict[int, int],
                          num_stages: int) -> Program:
    out: Program = {r: [] for r in ranks}
    last = num_stages - 1
    for row in timeline:
        for r, slot in zip(ranks, row, strict=True):
            if slot is None:
                continue
            out[r].append(slot)
            for act in flatten(slot):
                if act.kind == ActionKind.FORWARD and act.stage != last:
                    peer = stage_to_rank[act.stage + 1]
                    if peer != r:
                        out[r].append(Action(ActionKind.SEND_F, act.stage, act.microbatch))
                        out[peer].append(Action(ActionKind.RECV_F, act.stage + 1, act.microbatch))
                elif act.kind in (ActionKind.BACKWARD_FULL, ActionKind.BACKWARD_INPUT) and act.stage != 0:
                    peer = stage_to_rank[act.stage - 1]
                    if peer != r:
                        out[r].append(Action(ActionKind.SEND_B, act.stage, act.microbatch))
                        out[peer].append(Action(ActionKind.RECV_B, act.stage - 1, act.microbatch))
    return out


def validate_program(program: Program, stage_to_rank: dict[int, int], num_stages: int, num_microbatches: int,
                     has_backward: bool) -> None:
    """Replay the program with blocking semantics and check completeness; raises on deadlock / missing work."""
    cursor = {r: 0 for r in program}
    done: set[tuple[ActionKind, int, int]] = set()
    sent: set[tuple[ActionKind, int, int]] = set()
    last = num_stages - 1
    progressed = True
    while progressed:
        progressed = False
        for r, actions in program.items():
            while cursor[r] < len(actions):
                ok = True
                for act in flatten(actions[cursor[r]]):
                    k, s, m = act.kind, act.stage, act.microbatch
                    if k == ActionKind.FORWARD:
                        need = s == 0 or (ActionKind.FORWARD, s - 1, m) in done if stage_to_rank.get(s - 1) == r or s == 0 \
                            else (ActionKind.RECV_F, s, m) in done
                    elif k in (ActionKind.BACKWARD_FULL, ActionKind.BACKWARD_INPUT):
                        own = (ActionKind.FORWARD, s, m) in done
                        if s == last:
                            nxt = True
                        elif stage_to_rank[s + 1] == r:
                            nxt = (ActionKind.BACKWARD_FULL, s + 1, m) in done or (ActionKind.BACKWARD_INPUT, s + 1, m) in done
                        else:
                            nxt = (ActionKind.RECV_B, s, m) in done
                        need = own and nxt
                    elif k == ActionKind.BACKWARD_WEIGHT:
                        need = (ActionKind.BACKWARD_INPUT, s, m) in done
                    elif k == ActionKind.SEND_F:
                        need = (ActionKind.FORWARD, s, m) in done
                    elif k == ActionKind.SEND_B:
                        need = (ActionKind.BACKWARD_FULL, s, m) in done or (ActionKind.BACKWARD_INPUT, s, m) in done
                    elif k == ActionKind.RECV_F:
                        need = (ActionKind.SEND_F, s - 1, m) in sent
                    else:  # RECV_B
                        need = (ActionKind.SEND_B, s + 1, m) in sent
                    if not need:
                        ok = False
                        break
                    done.add((k, s, m))
                    if k in (ActionKind.SEND_F, ActionKind.SEND_B):
                        sent.add((k, s, m))
                if not ok:
                    break
                cursor[r] += 1
                progressed = True
    stuck = {r: str(a[cursor[r]]) for r, a in program.items() if cursor[r] < len(a)}
    if stuck:
        raise RuntimeError(f"Deadlock in pipeline program, ranks blocked at: {stuck}")
    for s in range(num_stages):
        for m in range(num_microbatches):
            if (ActionKind.FORWARD, s, m) not in done:
                raise RuntimeError(f"program misses forward of stage {s} microbatch {m}")
            if has_backward:
                full = (ActionKind.BACKWARD_FULL, s, m) in done
                split = (ActionKind.BACKWARD_INPUT, s, m) in done and (ActionKind.BACKWARD_WEIGHT, s, m) in done
                if not (full or split):
                    raise RuntimeError(f"program misses backward of stage {s} microbatch {m}")
