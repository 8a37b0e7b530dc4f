"""LR-scheduler provider driven by a discriminated pydantic config (``name`` selects the family)."""

from __future__ import annotations

from collections.abc import Callable
from typing import Annotated, Literal

from pydantic import BaseModel, Field

from d9d_b200.core.protocol import LRSchedulerProtocol
from d9d_b200.loop.control import InitializeLRSchedulerContext, LRSchedulerProvider
from d9d_b200.lr_scheduler.piecewise import PiecewiseSchedulerConfig, piecewise_scheduler_from_config


class PiecewiseConfig(BaseModel):
    name: Literal["piecewise"] = "piecewise"
    scheduler: PiecewiseSchedulerConfig


AutoLRSchedulerConfig = Annotated[PiecewiseConfig, Field(discriminator="name")]

_BUILDERS: dict[type, Callable[[BaseModel, InitializeLRSchedulerContext], LRSchedulerProtocol]] = {
    PiecewiseConfig: lambda cfg, ctx: piecewise_scheduler_from_config(cfg.scheduler, optimizer=ctx.optimizer, total_steps=ctx.total_steps),
}


class AutoLRSchedulerProvider(LRSchedulerProvider):
    def __init__(self, config: PiecewiseConfig):
        if type(config) not in _BUILDERS:
            raise ValueError(f"Unsupported LR scheduler type: {config}")
        self._config = config

    def __call__(self, context: InitializeLRSchedulerContext) -> LRSchedulerProtocol:
        return _BUILDERS[type(self._config)](self._config, context)
