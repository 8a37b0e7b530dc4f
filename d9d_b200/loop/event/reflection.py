from __future__ import annotations

import inspect
from collections.abc import Callable

from .core import Event, EventBus, TContext

_MARK = "_d9d_subscribed_events"


def subscribe(event: Event[TContext]) -> Callable[[Callable[[TContext], None]], Callable[[TContext], None]]:
    """Tag a method as a handler of ``event``; :func:`subscribe_annotated` performs the registration."""

    def decorator(func: Callable) -> Callable:
        setattr(func, _MARK, event)
        return func

    return decorator


def subscribe_annotated(bus: EventBus, target: object) -> None:
    """Register every ``@subscribe``-tagged bound method of ``target`` on ``bus``."""
    for _, method in inspect.getmembers(target, predicate=inspect.ismethod):
        event = getattr(method.__func__, _MARK, None)
        if event is not None:
            bus.subscribe(event, method)
