from __future__ import annotations

import abc
from collections.abc import Generator
from contextlib import contextmanager
from typing import Any, Generic, Self, TypeVar

import torch
from pydantic import BaseModel, Field
from torch.distributed.checkpoint.stateful import Stateful


class BaseTrackerRun(abc.ABC):
    """An open run: receives the current step, a persistent tag context, scalars and histograms."""

    @abc.abstractmethod
    def set_step(self, step: int) -> None: ...

    @abc.abstractmethod
    def set_context(self, context: dict[str, str]) -> None: ...

    @abc.abstractmethod
    def scalar(self, name: str, value: float, context: dict[str, str] | None = None) -> None: ...

    @abc.abstractmethod
    def bins(self, name: str, values: torch.Tensor, context: dict[str, str] | None = None) -> None: ...


class RunConfig(BaseModel):
    name: str
    description: str | None
    hparams: dict[str, Any] = Field(default_factory=dict)


TConfig = TypeVar("TConfig", bound=BaseModel)


class BaseTracker(abc.ABC, Stateful, Generic[TConfig]):
    """Backend factory; Stateful so that a resumed job continues the same run."""

    @contextmanager
    @abc.abstractmethod
    def open(self, properties: RunConfig) -> Generator[BaseTrackerRun, None, None]: ...

    @classmethod
    @abc.abstractmethod
    def from_config(cls, config: TConfig) -> Self: ...
