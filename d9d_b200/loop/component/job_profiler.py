from __future__ import annotations

from collections.abc import Generator
from contextlib import contextmanager

import torch.profiler

from d9d_b200.core.dist_context import DistributedContext
from d9d_b200.internals.profiling import Profiler
from d9d_b200.loop.config import ProfilingConfig

from .stepper import Stepper


class JobProfiler:
    """Optional profiler (``config.enabled``) resuming its schedule from the current step."""

    def __init__(self, dist_context: DistributedContext, config: ProfilingConfig | None, stepper: Stepper):
        self._stepper = stepper
        self._profiler = None
        if config is not None and config.enabled:
            self._profiler = Profiler(save_dir=config.traces_dir, period_steps=config.period_steps, warmup_steps=config.warmup_steps,
                                      active_steps=config.active_steps, dist_context=dist_context)

    @contextmanager
    def open(self) -> Generator[torch.profiler.profile | None, None, None]:
        if self._profiler is None:
            yield None
        else:
            with self._profiler.open(self._stepper.current_step) as prof:
                yield prof
