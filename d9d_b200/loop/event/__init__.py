"""Typed event bus (reference ``d9d/loop/event``)."""

from .core import Event, EventBus, TContext
from .reflection import subscribe, subscribe_annotated

__all__ = ["Event", "EventBus", "TContext", "subscribe", "subscribe_annotated"]
