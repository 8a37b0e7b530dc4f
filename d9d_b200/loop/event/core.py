from __future__ import annotations

import contextlib
import dataclasses
from collections.abc import Callable, Generator
from typing import Any, Generic, TypeVar

TContext = TypeVar("TContext")


@dataclasses.dataclass(frozen=True)
class Event(Generic[TContext]):
    """Event descriptor: a unique id; the generic parameter documents the context type handlers receive."""

    id: str


class EventBus:
    """Synchronous publish/subscribe: handlers run in subscription order on ``trigger``."""

    def __init__(self) -> None:
        self._handlers: dict[Event, list[Callable[[Any], None]]] = {}

    def subscribe(self, event: Event[TContext], handler: Callable[[TContext], None]) -> None:
        self._handlers.setdefault(event, []).append(handler)

    def trigger(self, event: Event[TContext], context: TContext) -> None:
        for handler in self._handlers.get(event, ()):
            handler(context)

    @contextlib.contextmanager
    def bounded(self, event_pre: Event[TContext], event_post: Event[TContext], context: TContext) -> Generator[None, None, None]:
        """``event_pre`` on entry, ``event_post`` after the block completed successfully."""
        self.trigger(event_pre, context)
        yield
        self.trigger(event_post, context)
