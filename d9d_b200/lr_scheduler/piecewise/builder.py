from __future__ import annotations

from typing import Self

from torch.optim import Optimizer
from torch.optim.lr_scheduler import LambdaLR

from d9d_b200.core.protocol import LRSchedulerProtocol

from .curves import CurveBase
from .engine import PiecewiseScheduleEngine, SchedulePhase


class PiecewiseScheduleBuilder:
    """Fluent definition of consecutive phases: ``for_steps`` / ``until_percentage`` / ``fill_rest`` then ``build``.

    Parity: reference ``d9d/lr_scheduler/piecewise/builder.py:12-141``.
    """

    def __init__(self, initial_multiplier: float, total_steps: int | None):
        self._phases: list[SchedulePhase] = []
        self._total_steps = total_steps
        self._cursor = 0
        self._value = initial_multiplier

    def for_steps(self, steps: int, target_multiplier: float, curve: CurveBase) -> Self:
        self._phases.append(SchedulePhase(self._cursor, self._cursor + steps, self._value, target_multiplier, curve))
        self._cursor += steps
        self._value = target_multiplier
        return self

    def until_percentage(self, p: float, target_multiplier: float, curve: CurveBase) -> Self:
        if self._total_steps is None:
            raise ValueError("You must define 'total_steps' in the constructor to use percentage-based methods.")
        if not 0.0 <= p <= 1.0:
            raise ValueError("Percentage should be in range of [0.0, 1.0]")
        target_step = int(self._total_steps * p)
        if target_step < self._cursor:
            raise ValueError(f"Target percentage {p} (step {target_step}) is behind current cursor (step {self._cursor}).")
        return self.for_steps(target_step - self._cursor, target_multiplier, curve)

    def fill_rest(self, target_multiplier: float, curve: CurveBase) -> Self:
        return self.until_percentage(1.0, target_multiplier, curve)

    def build(self, optimizer: Optimizer) -> LRSchedulerProtocol:
        if self._total_steps is not None and self._cursor > self._total_steps:
            raise ValueError(f"Schedule defined for {self._cursor} steps, but total_steps is {self._total_steps}.")
        return LambdaLR(optimizer, PiecewiseScheduleEngine(self._phases).get_factor)


def piecewise_schedule(initial_multiplier: float, total_steps: int | None = None) -> PiecewiseScheduleBuilder:
    return PiecewiseScheduleBuilder(initial_multiplier=initial_multiplier, total_steps=total_steps)
