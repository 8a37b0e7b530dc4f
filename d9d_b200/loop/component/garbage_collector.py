from __future__ import annotations

import gc
import time
from contextlib import AbstractContextManager
from types import TracebackType
from typing import Self

from d9d_b200.core.dist_context import DistributedContext
from d9d_b200.loop.config import GarbageCollectionConfig

from .stepper import Stepper


class ManualGarbageCollector(AbstractContextManager):
    """Disables automatic GC inside the loop (no random pauses desynchronising ranks); collects generation 1
    periodically and generation 2 on demand / on enter / on exit.  Pause durations are logged."""

    def __init__(self, dist_ctx: DistributedContext, config: GarbageCollectionConfig, step: Stepper):
        self._ctx, self._config, self._step = dist_ctx, config, step

    def __enter__(self) -> Self:
        gc.disable()
        self._collect(2)
        return self

    def __exit__(self, exc_type: type[BaseException] | None, exc_value: BaseException | None, traceback: TracebackType | None, /) -> None:
        gc.enable()
        self._collect(2)

    def collect_periodic(self) -> None:
        if self._step.should_do_action(self._config.period_steps, enable_on_last_step_if_periodic=False):
            self._collect(1)

    def collect_forced(self) -> None:
        self._collect(2)

    def _collect(self, generation: int) -> None:
        t0 = time.monotonic()
        gc.collect(generation)
        self._ctx.logger.info(f"[GC] Garbage collection for generation {generation} took {time.monotonic() - t0}s")
