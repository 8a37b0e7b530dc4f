"""Pieces of job assembly shared by the training and the inference configurators."""

from __future__ import annotations

import dataclasses
from collections.abc import Callable
from typing import Any

from d9d_b200.core.dist_context import DeviceMeshParameters, DistributedContext
from d9d_b200.internals.determinism import set_seeds
from d9d_b200.loop import component as parts
from d9d_b200.loop import control
from d9d_b200.loop.config import BatchingConfig, PipeliningConfig
from d9d_b200.loop.event import Event, EventBus
from d9d_b200.loop.event.catalogue import common as shared_events


@dataclasses.dataclass
class JobFoundation:
    """What exists before any data or model is touched: the distributed context (seeded), the armed init timeout,
    the user's task with its (and the model provider's) event subscriptions, and the batch arithmetic."""

    ctx: DistributedContext
    timeout: parts.TimeoutManager
    task: Any
    bus: EventBus
    maths: parts.BatchMaths


def lay_foundation(mesh: DeviceMeshParameters, config: Any, make_task: Callable[[DistributedContext], Any],
                   model_provider: control.ModelProvider, batching: BatchingConfig, pipelining: PipeliningConfig | None,
                   started: Event) -> JobFoundation:
    """Build the context, seed every RNG, arm the long init timeout, wire events and announce ``started``."""
    ctx = mesh.build()
    set_seeds(ctx, seed=config.determinism.base_seed)
    timeout = parts.TimeoutManager(dist_context=ctx, config=config.timeout)
    timeout.set_init()
    task = make_task(ctx)
    bus = EventBus()
    model_provider.register_events(control.RegisterModelEventsContext(dist_context=ctx, event_bus=bus))
    task.register_events(control.RegisterTaskEventsContext(dist_context=ctx, event_bus=bus))
    bus.trigger(started, shared_events.EventConfigurationStartedContext(dist_context=ctx))
    maths = parts.BatchMaths(dist_context=ctx, config_batching=batching, config_pipelining=pipelining)
    return JobFoundation(ctx=ctx, timeout=timeout, task=task, bus=bus, maths=maths)


@dataclasses.dataclass
class Housekeeping:
    gc: parts.ManualGarbageCollector
    checkpointer: parts.StateCheckpointer
    profiler: parts.JobProfiler


def build_housekeeping(ctx: DistributedContext, config: Any, stepper: parts.Stepper, run_name: str | None) -> Housekeeping:
    """Manual GC, job checkpointing and the periodic profiler - all driven by the same stepper."""
    gc = parts.ManualGarbageCollector(dist_ctx=ctx, config=config.gc, step=stepper)
    return Housekeeping(
        gc=gc,
        checkpointer=parts.StateCheckpointer(dist_context=ctx, stepper=stepper, config=config.checkpointing, gc=gc, run_name=run_name),
        profiler=parts.JobProfiler(dist_context=ctx, stepper=stepper, config=config.profiling),
    )
