from typing import Annotated, Literal

from pydantic import BaseModel, Field


class PipelineScheduleInferenceConfig(BaseModel):
    """Forward-only execution."""

    schedule: Literal["inference"] = "inference"


class PipelineScheduleGPipeConfig(BaseModel):
    """All forwards, then all backwards; one stage per rank."""

    schedule: Literal["gpipe"] = "gpipe"


class PipelineScheduleLoopedBFSConfig(BaseModel):
    """GPipe generalised to several (looped) stages per rank."""

    schedule: Literal["looped_bfs"] = "looped_bfs"
    num_stages_per_rank: int


class PipelineSchedule1F1BConfig(BaseModel):
    """(Interleaved) 1F1B; ``zero_bubble`` splits the backward into dI and dW (ZB1P)."""

    schedule: Literal["1f1b"] = "1f1b"
    num_stages_per_rank: int
    zero_bubble: bool


class PipelineScheduleZeroBubbleVConfig(BaseModel):
    """Zero-bubble on the V topology (exactly two stages per rank)."""

    schedule: Literal["zero_bubble_v"] = "zero_bubble_v"


class PipelineScheduleDualPipeVConfig(BaseModel):
    """DualPipeV: V topology with paired forward/backward slots."""

    schedule: Literal["dual_pipe_v"] = "dual_pipe_v"


AnyPipelineScheduleConfig = Annotated[
    PipelineScheduleInferenceConfig
    | PipelineScheduleGPipeConfig
    | PipelineScheduleLoopedBFSConfig
    | PipelineSchedule1F1BConfig
    | PipelineScheduleZeroBubbleVConfig
    | PipelineScheduleDualPipeVConfig,
    Field(discriminator="schedule"),
]
