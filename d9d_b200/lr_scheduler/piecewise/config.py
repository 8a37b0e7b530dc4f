from __future__ import annotations

from typing import Annotated, Literal

from pydantic import BaseModel, Field, PositiveInt
from torch.optim import Optimizer

from d9d_b200.core.protocol import LRSchedulerProtocol

from .builder import piecewise_schedule
from .curves import CurveBase, CurveCosine, CurveExponential, CurveLinear, CurvePoly


class CurveLinearConfig(BaseModel):
    type: Literal["linear"] = "linear"


class CurveCosineConfig(BaseModel):
    type: Literal["cosine"] = "cosine"


class CurveExponentialConfig(BaseModel):
    type: Literal["exponential"] = "exponential"


class CurvePolyConfig(BaseModel):
    type: Literal["poly"] = "poly"
    power: float = 2.0


AnyCurveConfig = Annotated[CurveLinearConfig | CurveCosineConfig | CurveExponentialConfig | CurvePolyConfig, Field(discriminator="type")]


def curve_from_config(config: CurveLinearConfig | CurveCosineConfig | CurveExponentialConfig | CurvePolyConfig) -> CurveBase:
    if isinstance(config, CurveLinearConfig):
        return CurveLinear()
    if isinstance(config, CurveCosineConfig):
        return CurveCosine()
    if isinstance(config, CurveExponentialConfig):
        return CurveExponential()
    if isinstance(config, CurvePolyConfig):
        return CurvePoly(config.power)
    raise TypeError(f"unknown curve config {type(config).__name__}")


class StepPhaseConfig(BaseModel):
    mode: Literal["steps"] = "steps"
    steps: PositiveInt
    target_multiplier: float
    curve: AnyCurveConfig


class PercentagePhaseConfig(BaseModel):
    mode: Literal["percentage"] = "percentage"
    percentage: float = Field(..., ge=0.0, le=1.0)
    target_multiplier: float
    curve: AnyCurveConfig


class RestPhaseConfig(BaseModel):
    mode: Literal["rest"] = "rest"
    target_multiplier: float
    curve: AnyCurveConfig


PhaseConfig = Annotated[StepPhaseConfig | PercentagePhaseConfig | RestPhaseConfig, Field(discriminator="mode")]


class PiecewiseSchedulerConfig(BaseModel):
    initial_multiplier: float
    phases: list[PhaseConfig]


def piecewise_scheduler_from_config(config: PiecewiseSchedulerConfig, optimizer: Optimizer, total_steps: int | None) -> LRSchedulerProtocol:
    builder = piecewise_schedule(config.initial_multiplier, total_steps)
    for phase in config.phases:
        curve = curve_from_config(phase.curve)
        if isinstance(phase, StepPhaseConfig):
            builder.for_steps(phase.steps, phase.target_multiplier, curve)
        elif isinstance(phase, PercentagePhaseConfig):
            builder.until_percentage(phase.percentage, phase.target_multiplier, curve)
        else:
            builder.fill_rest(phase.target_multiplier, curve)
    return builder.build(optimizer)
