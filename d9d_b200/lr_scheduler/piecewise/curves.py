from __future__ import annotations

import abc
import math


class CurveBase(abc.ABC):
    """Interpolation between the multiplier at the start and at the end of a phase, ``step_p`` in [0, 1]."""

    @abc.abstractmethod
    def compute(self, start: float, end: float, step_p: float) -> float: ...


class CurveLinear(CurveBase):
    def compute(self, start: float, end: float, step_p: float) -> float:
        return start + (end - start) * step_p


class CurveCosine(CurveBase):
    """Half-period cosine from ``start`` to ``end``."""

    def compute(self, start: float, end: float, step_p: float) -> float:
        return end + (start - end) * 0.5 * (1.0 + math.cos(math.pi * step_p))


class CurvePoly(CurveBase):
    def __init__(self, power: float):
        self._power = power

    def compute(self, start: float, end: float, step_p: float) -> float:
        return start + (end - start) * step_p**self._power


class CurveExponential(CurveBase):
    """Linear in log-space (values are floored at 1e-8)."""

    def compute(self, start: float, end: float, step_p: float) -> float:
        lo, hi = math.log(max(start, 1e-8)), math.log(max(end, 1e-8))
        return math.exp(lo + (hi - lo) * step_p)
