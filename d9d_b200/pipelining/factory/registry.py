from __future__ import annotations

from collections.abc import Callable
from typing import Any

from d9d_b200.pipelining.infra.programs import (
    DualPipeVPipelineProgramBuilder,
    Interleaved1F1BPipelineProgramBuilder,
    LoopedBFSPipelineProgramBuilder,
    PipelineProgramBuilder,
    ZeroBubbleVPipelineProgramBuilder,
)

from .config import (
    PipelineSchedule1F1BConfig,
    PipelineScheduleDualPipeVConfig,
    PipelineScheduleGPipeConfig,
    PipelineScheduleInferenceConfig,
    PipelineScheduleLoopedBFSConfig,
    PipelineScheduleZeroBubbleVConfig,
)


class PipelineProgramRegistry:
    """config class -> program-builder factory; user schedules can be registered with the decorator."""

    def __init__(self) -> None:
        self._factories: dict[type, Callable[[Any], PipelineProgramBuilder]] = {}

    def register_program(self, config_cls: type) -> Callable[[Callable[[Any], PipelineProgramBuilder]], Callable[[Any], PipelineProgramBuilder]]:
        def decorator(fn: Callable[[Any], PipelineProgramBuilder]) -> Callable[[Any], PipelineProgramBuilder]:
            self._factories[config_cls] = fn
            return fn

        return decorator

    def program_for(self, config: Any) -> PipelineProgramBuilder:
        try:
            return self._factories[type(config)](config)
        except KeyError:
            raise ValueError(f"no pipeline program registered for {type(config).__name__}") from None


PIPELINE_PROGRAM_REGISTRY = PipelineProgramRegistry()
PIPELINE_PROGRAM_REGISTRY.register_program(PipelineScheduleGPipeConfig)(lambda _: LoopedBFSPipelineProgramBuilder(1, inference_mode=False))
PIPELINE_PROGRAM_REGISTRY.register_program(PipelineScheduleInferenceConfig)(lambda _: LoopedBFSPipelineProgramBuilder(1, inference_mode=True))
PIPELINE_PROGRAM_REGISTRY.register_program(PipelineScheduleLoopedBFSConfig)(lambda c: LoopedBFSPipelineProgramBuilder(c.num_stages_per_rank, inference_mode=False))
PIPELINE_PROGRAM_REGISTRY.register_program(PipelineSchedule1F1BConfig)(lambda c: Interleaved1F1BPipelineProgramBuilder(c.num_stages_per_rank, enable_zero_bubble=c.zero_bubble))
PIPELINE_PROGRAM_REGISTRY.register_program(PipelineScheduleZeroBubbleVConfig)(lambda _: ZeroBubbleVPipelineProgramBuilder())
PIPELINE_PROGRAM_REGISTRY.register_program(PipelineScheduleDualPipeVConfig)(lambda _: DualPipeVPipelineProgramBuilder())
