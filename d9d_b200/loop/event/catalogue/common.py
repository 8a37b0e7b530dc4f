"""Payloads shared by both loops (defined in ``contexts.py``) and the helper that mints namespaced events."""

from __future__ import annotations

from typing import TypeVar

from d9d_b200.loop.event import Event

from .contexts import (
    EventConfigurationStartedContext,
    EventDataLoaderReadyContext,
    EventModelStagesReadyContext,
    EventStepContext,
)

T = TypeVar("T")


def namespaced(namespace: str):
    """``make = namespaced("train"); make("step.pre", EventStepContext)`` -> ``Event[EventStepContext]("train.step.pre")``."""

    def make(name: str, payload: type[T]) -> Event[T]:
        return Event[payload](id=f"{namespace}.{name}")  # type: ignore[valid-type]

    return make


__all__ = [
    "EventConfigurationStartedContext",
    "EventDataLoaderReadyContext",
    "EventModelStagesReadyContext",
    "EventStepContext",
    "namespaced",
]
