"""Tracker that drops everything (non-main ranks, unit tests, benchmarks)."""

from __future__ import annotations

from collections.abc import Iterator
from contextlib import contextmanager
from typing import Any, Literal, Self

from pydantic import BaseModel

from d9d_b200.tracker.base import BaseTracker, BaseTrackerRun, RunConfig


class NullTrackerConfig(BaseModel):
    provider: Literal["null"] = "null"


class NullRun(BaseTrackerRun):
    def _discard(self, *args: Any, **kwargs: Any) -> None:
        return None

    set_step = set_context = scalar = bins = _discard  # type: ignore[assignment]


_THE_RUN = NullRun()


class NullTracker(BaseTracker[NullTrackerConfig]):
    @classmethod
    def from_config(cls, config: NullTrackerConfig) -> Self:
        return cls()

    @contextmanager
    def open(self, properties: RunConfig) -> Iterator[BaseTrackerRun]:
        yield _THE_RUN

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:
        del state_dict

    def state_dict(self) -> dict[str, Any]:
        return {}
