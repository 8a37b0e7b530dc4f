from __future__ import annotations

from typing import Annotated, Literal

from pydantic import BaseModel, Field

from d9d_b200.core.protocol import LRSchedulerProtocol
from d9d_b200.loop.control import InitializeLRSchedulerContext, LRSchedulerProvider
from d9d_b200.lr_scheduler.piecewise import PiecewiseSchedulerConfig, piecewise_scheduler_from_config


class PiecewiseConfig(BaseModel):
    name: Literal["piecewise"] = "piecewise"
    scheduler: PiecewiseSchedulerConfig


AutoLRSchedulerConfig = Annotated[PiecewiseConfig, Field(discriminator="name")]


class AutoLRSchedulerProvider(LRSchedulerProvider):
    def __init__(self, config: PiecewiseConfig):
        self._config = config

    def __call__(self, context: InitializeLRSchedulerContext) -> LRSchedulerProtocol:
        if isinstance(self._config, PiecewiseConfig):
            return piecewise_scheduler_from_config(self._config.scheduler, optimizer=context.optimizer, total_steps=context.total_steps)
        raise ValueError(f"Unsupported LR scheduler type: {self._config}")
