"""Action IR of a pipeline program: every rank executes a flat list of these.

Names follow the reference's trace labels (``3F1``, ``2SEND_B0``, ``5I2|4F3`` …; reference
``d9d/pipelining/infra/schedule/component/runtime/action.py``) so profiler traces stay comparable.
"""

from __future__ import annotations

import dataclasses
import enum


class ActionKind(enum.StrEnum):
    FORWARD = "F"
    BACKWARD_FULL = "B"  # dI and dW together
    BACKWARD_INPUT = "I"  # dI only (dW deferred)
    BACKWARD_WEIGHT = "W"  # deferred dW
    SEND_F = "SEND_F"
    RECV_F = "RECV_F"
    SEND_B = "SEND_B"
    RECV_B = "RECV_B"


_COMPUTE = {ActionKind.FORWARD, ActionKind.BACKWARD_FULL, ActionKind.BACKWARD_INPUT, ActionKind.BACKWARD_WEIGHT}
_BACKWARD = {ActionKind.BACKWARD_FULL, ActionKind.BACKWARD_INPUT, ActionKind.BACKWARD_WEIGHT, ActionKind.SEND_B,
             ActionKind.RECV_B}


@dataclasses.dataclass(frozen=True, slots=True)
class Action:
    kind: ActionKind
    stage: int
    microbatch: int

    @property
    def is_compute(self) -> bool:
        return self.kind in _COMPUTE

    @property
    def has_backward_work(self) -> bool:
        return self.kind in _BACKWARD

    @property
    def stage_idx(self) -> int:
        return self.stage

    @property
    def microbatch_idx(self) -> int:
        return self.microbatch

    def __str__(self) -> str:
        return f"{self.stage}{self.kind.value}{self.microbatch}"


@dataclasses.dataclass(frozen=True, slots=True)
class ComposeAction:
    """Several compute actions issued back to back in one schedule slot (forward/backward pairing of DualPipeV)."""

    actions: tuple[Action, ...]

    @property
    def is_compute(self) -> bool:
        return True

    @property
    def has_backward_work(self) -> bool:
        return any(a.has_backward_work for a in self.actions)

    def __str__(self) -> str:
        return "|".join(str(a) for a in self.actions)


AnyAction = Action | ComposeAction
Program = dict[int, list[AnyAction]]


def F(stage: int, mb: int) -> Action:  # noqa: N802
    return Action(ActionKind.FORWARD, stage, mb)


def B(stage: int, mb: int) -> Action:  # noqa: N802
    return Action(ActionKind.BACKWARD_FULL, stage, mb)


def I(stage: int, mb: int) -> Action:  # noqa: N802, E743
    return Action(ActionKind.BACKWARD_INPUT, stage, mb)


def W(stage: int, mb: int) -> Action:  # noqa: N802
    return Action(ActionKind.BACKWARD_WEIGHT, stage, mb)


def flatten(action: AnyAction) -> tuple[Action, ...]:
    return action.actions if isinstance(action, ComposeAction) else (action,)
