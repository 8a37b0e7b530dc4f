"""Piecewise multi-phase LR schedules built with a fluent builder or from a pydantic config."""

from .builder import PiecewiseScheduleBuilder, piecewise_schedule
from .config import PiecewiseSchedulerConfig, piecewise_scheduler_from_config
from .curves import CurveBase, CurveCosine, CurveExponential, CurveLinear, CurvePoly

__all__ = [
    "CurveBase",
    "CurveCosine",
    "CurveExponential",
    "CurveLinear",
    "CurvePoly",
    "PiecewiseScheduleBuilder",
    "PiecewiseSchedulerConfig",
    "piecewise_schedule",
    "piecewise_scheduler_from_config",
]
