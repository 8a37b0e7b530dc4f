"""Pipeline runtime (reference ``d9d/pipelining/infra/schedule/component/runtime/__init__.py``).

The reference models every action as its own class; the IR here is one frozen ``Action(kind, stage, microbatch)``
record (cheap to hash / compare - the simulator creates and matches many of them).  The class names below are
*views* over that record: calling one builds the record, ``isinstance`` checks the kind.
"""

from __future__ import annotations

from d9d_b200.pipelining.infra.action import Action, ActionKind, ComposeAction
from d9d_b200.pipelining.infra.executor import OfflinePipelineExecutor, PipelineScheduleExecutor

ActionBase = Action | ComposeAction


class _ActionView(type):
    kinds: tuple[ActionKind, ...] = ()

    def __call__(cls, stage_idx: int, microbatch_idx: int) -> Action:  # type: ignore[override]
        return Action(cls.kinds[0], stage_idx, microbatch_idx)

    def __instancecheck__(cls, obj: object) -> bool:
        return isinstance(obj, Action) and obj.kind in cls.kinds


class ForwardComputeAction(metaclass=_ActionView):
    kinds = (ActionKind.FORWARD,)


class BackwardWeightComputeAction(metaclass=_ActionView):
    kinds = (ActionKind.BACKWARD_WEIGHT,)


class ForwardSendAction(metaclass=_ActionView):
    kinds = (ActionKind.SEND_F,)


class ForwardReceiveAction(metaclass=_ActionView):
    kinds = (ActionKind.RECV_F,)


class BackwardSendAction(metaclass=_ActionView):
    kinds = (ActionKind.SEND_B,)


class BackwardReceiveAction(metaclass=_ActionView):
    kinds = (ActionKind.RECV_B,)


class _BackwardView(_ActionView):
    def __call__(cls, stage_idx: int, microbatch_idx: int, full_backward: bool) -> Action:  # type: ignore[override]
        kind = ActionKind.BACKWARD_FULL if full_backward else ActionKind.BACKWARD_INPUT
        return Action(kind, stage_idx, microbatch_idx)


class BackwardFullInputComputeAction(metaclass=_BackwardView):
    """Input-gradient backward; with ``full_backward=True`` the weight gradients are computed in the same action."""

    kinds = (ActionKind.BACKWARD_FULL, ActionKind.BACKWARD_INPUT)


__all__ = [
    "ActionBase",
    "BackwardFullInputComputeAction",
    "BackwardReceiveAction",
    "BackwardSendAction",
    "BackwardWeightComputeAction",
    "ComposeAction",
    "ForwardComputeAction",
    "ForwardReceiveAction",
    "ForwardSendAction",
    "OfflinePipelineExecutor",
    "PipelineScheduleExecutor",
]
