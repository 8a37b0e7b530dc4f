"""Schedule configuration and construction (reference ``d9d/pipelining/factory``)."""

from .config import (
    AnyPipelineScheduleConfig,
    PipelineSchedule1F1BConfig,
    PipelineScheduleDualPipeVConfig,
    PipelineScheduleGPipeConfig,
    PipelineScheduleInferenceConfig,
    PipelineScheduleLoopedBFSConfig,
    PipelineScheduleZeroBubbleVConfig,
)
from .factory import PipelineScheduleInfo, build_schedule
from .registry import PIPELINE_PROGRAM_REGISTRY

__all__ = [
    "PIPELINE_PROGRAM_REGISTRY",
    "AnyPipelineScheduleConfig",
    "PipelineSchedule1F1BConfig",
    "PipelineScheduleDualPipeVConfig",
    "PipelineScheduleGPipeConfig",
    "PipelineScheduleInferenceConfig",
    "PipelineScheduleInfo",
    "PipelineScheduleLoopedBFSConfig",
    "PipelineScheduleZeroBubbleVConfig",
    "build_schedule",
]
