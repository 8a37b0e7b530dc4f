from __future__ import annotations

from collections.abc import Generator
from contextlib import contextmanager
from typing import Any, Literal, Self

import torch
from pydantic import BaseModel

from d9d_b200.tracker.base import BaseTracker, BaseTrackerRun, RunConfig


class NullTrackerConfig(BaseModel):
    provider: Literal["null"] = "null"


class NullRun(BaseTrackerRun):
    def set_step(self, step: int) -> None: ...

    def set_context(self, context: dict[str, str]) -> None: ...

    def scalar(self, name: str, value: float, context: dict[str, str] | None = None) -> None: ...

    def bins(self, name: str, values: torch.Tensor, context: dict[str, str] | None = None) -> None: ...


class NullTracker(BaseTracker[NullTrackerConfig]):
    """Discards everything (non-main ranks, tests)."""

    @contextmanager
    def open(self, properties: RunConfig) -> Generator[BaseTrackerRun, None, None]:
        yield NullRun()

    @classmethod
    def from_config(cls, config: NullTrackerConfig) -> Self:
        return cls()

    def state_dict(self) -> dict[str, Any]:
        return {}

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:
        return None
