"""Throughput / model-FLOPs-utilisation meter as an event-bus plug-in (the reference has no throughput meter)."""

from __future__ import annotations

import time

import torch

from d9d_b200.loop.event import EventBus, subscribe, subscribe_annotated
from d9d_b200.loop.event.catalogue.train import (
    EVENT_TRAIN_READY,
    EVENT_TRAIN_STEP_POST,
    EVENT_TRAIN_STEP_PRE,
    EventStepContext,
    EventTrainReadyContext,
)
from d9d_b200.tracker import BaseTrackerRun


def transformer_flops_per_token(num_parameters_active: int, num_layers: int, hidden_size: int, seq_len: int,
                                causal: bool = True) -> float:
    """Training FLOPs per token: ``6 * N_active`` for the matmuls with weights plus attention scores / values
    (``12 * L * H * S`` forward + backward, halved for causal masking)."""
    attention = 12.0 * num_layers * hidden_size * seq_len
    return 6.0 * num_parameters_active + (attention / 2 if causal else attention)


class ThroughputMeter:
    """Logs ``tokens_per_second`` (whole job) and, given ``flops_per_token`` and ``peak_flops_per_device``, ``mfu`` every
    ``period_steps`` optimizer steps.

    Steps are timed on the device with CUDA events recorded at ``step.pre`` / ``step.post``; an interval is only read
    back once a *later* step has been recorded, so the meter never makes the host wait for the GPU.  On CPU it uses the
    wall clock.  Register with ``meter.install(event_bus)`` from ``ModelProvider.register_events`` or
    ``Task.register_events``.
    """

    def __init__(self, tokens_per_step: int, world_size: int = 1, flops_per_token: float | None = None,
                 peak_flops_per_device: float | None = None, period_steps: int = 10, skip_first_steps: int = 1) -> None:
        self._tokens = tokens_per_step
        self._world = world_size
        self._flops_per_token = flops_per_token
        self._peak = peak_flops_per_device
        self._period = max(1, period_steps)
        self._skip = skip_first_steps
        self._run: BaseTrackerRun | None = None
        self._open: tuple[object, object] | None = None
        self._closed: list[tuple[object, object]] = []
        self._seen = 0
        self.history: list[dict[str, float]] = []

    def install(self, bus: EventBus) -> None:
        subscribe_annotated(bus, self)

    @staticmethod
    def _now() -> object:
        if torch.cuda.is_available():
            event = torch.cuda.Event(enable_timing=True)
            event.record()
            return event
        return time.perf_counter()

    @staticmethod
    def _seconds(start: object, end: object) -> float:
        if isinstance(start, float):
            return float(end) - start  # type: ignore[arg-type]
        return start.elapsed_time(end) / 1e3  # type: ignore[attr-defined]

    @subscribe(EVENT_TRAIN_READY)
    def _on_ready(self, ctx: EventTrainReadyContext) -> None:
        self._run = ctx.run

    @subscribe(EVENT_TRAIN_STEP_PRE)
    def _on_step_start(self, ctx: EventStepContext) -> None:
        self._open = (self._now(), None)

    @subscribe(EVENT_TRAIN_STEP_POST)
    def _on_step_end(self, ctx: EventStepContext) -> None:
        if self._open is None:
            return
        self._seen += 1
        if self._seen > self._skip:  # the first step(s) include lazy initialisation
            self._closed.append((self._open[0], self._now()))
        self._open = None
        # report the intervals recorded before the most recent one: their events have certainly completed
        if len(self._closed) > self._period:
            done, self._closed = self._closed[:-1], self._closed[-1:]
            self._report(done)

    def _report(self, intervals: list[tuple[object, object]]) -> None:
        seconds = sum(self._seconds(a, b) for a, b in intervals) / len(intervals)
        if seconds <= 0:
            return
        entry = {"tokens_per_second": self._tokens / seconds, "step_seconds": seconds}
        if self._flops_per_token is not None and self._peak is not None:
            entry["mfu"] = self._tokens * self._flops_per_token / seconds / (self._peak * self._world)
        self.history.append(entry)
        if self._run is not None:
            for name, value in entry.items():
                self._run.scalar(f"throughput/{name}", value)

    def flush(self) -> None:
        """Report whatever is still pending (synchronises with the device; call after training)."""
        if self._closed:
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            done, self._closed = self._closed, []
            self._report(done)
