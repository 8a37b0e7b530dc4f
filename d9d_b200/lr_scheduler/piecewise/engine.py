from __future__ import annotations

import bisect
import dataclasses

from .curves import CurveBase


@dataclasses.dataclass
class SchedulePhase:
    start_step: int
    end_step: int
    start_value: float
    end_value: float
    curve: CurveBase


class PiecewiseScheduleEngine:
    """Maps a global step to the LR multiplier; outside the defined phases the nearest boundary value holds."""

    def __init__(self, phases: list[SchedulePhase]):
        if not phases:
            raise ValueError("Scheduler should contain at least one phase")
        self._phases = [p for p in phases if p.end_step > p.start_step] or phases[:1]
        self._first, self._last = phases[0], phases[-1]
        self._starts = [p.start_step for p in self._phases]

    def get_factor(self, step: int) -> float:
        if step < self._first.start_step:
            return self._first.start_value
        if step >= self._last.end_step:
            return self._last.end_value
        phase = self._phases[max(bisect.bisect_right(self._starts, step) - 1, 0)]
        if not (phase.start_step <= step < phase.end_step):
            return self._last.end_value
        progress = (step - phase.start_step) / (phase.end_step - phase.start_step)
        return phase.curve.compute(start=phase.start_value, end=phase.end_value, step_p=progress)
