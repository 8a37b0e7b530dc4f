from __future__ import annotations

import enum
from collections.abc import Iterator
from contextlib import contextmanager


class GradDirection(enum.StrEnum):
    """Which family of gradient edges an operand belongs to."""

    inputs = "inputs"  # activations flowing to the previous layer / stage
    weight = "weights"  # parameters


_ALL = frozenset((GradDirection.inputs, GradDirection.weight))


class GlobalGradContext:
    """Holds the set of currently enabled gradient directions (both by default)."""

    __slots__ = ("_enabled",)

    def __init__(self) -> None:
        self._enabled: frozenset[GradDirection] = _ALL

    def check_direction(self, direction: GradDirection | None) -> bool:
        """``None`` (un-tagged operand) is always computed."""
        return direction is None or direction in self._enabled

    @property
    def enabled_directions(self) -> frozenset[GradDirection]:
        return self._enabled

    @contextmanager
    def with_directions(self, *directions: GradDirection) -> Iterator[None]:
        previous = self._enabled
        self._enabled = frozenset(directions)
        try:
            yield
        finally:
            self._enabled = previous


GLOBAL_GRAD_CONTEXT = GlobalGradContext()
