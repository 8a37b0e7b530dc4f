from __future__ import annotations

import enum

from d9d_b200.core.dist_context import DistributedContext
from d9d_b200.loop.config.config import TimeoutConfig


class TimeoutState(enum.StrEnum):
    none = "none"
    set_initial = "set_initial"
    set_regular = "set_regular"


class TimeoutManager:
    """Long collective timeout while the job initialises (first-step compilation, checkpoint loading), short one once
    steps are flowing — a hung rank is then detected within ``step_timeout`` seconds and the job can restart from
    the last checkpoint."""

    def __init__(self, dist_context: DistributedContext, config: TimeoutConfig):
        self._ctx, self._config = dist_context, config
        self._state = TimeoutState.none

    def set_init(self) -> None:
        if self._state != TimeoutState.none:
            raise ValueError("Can only set init timeout from initial state")
        self._ctx.set_timeout(self._config.init_timeout)
        self._state = TimeoutState.set_initial

    def set_periodic(self) -> None:
        if self._state == TimeoutState.set_initial:
            self._ctx.set_timeout(self._config.step_timeout)
            self._state = TimeoutState.set_regular
        elif self._state != TimeoutState.set_regular:
            raise ValueError("Unknown timeout state")
