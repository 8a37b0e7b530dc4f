"""Pipeline schedule selection (pydantic, discriminated by ``schedule``)."""

from typing import Annotated, ClassVar, Literal

from pydantic import BaseModel, Field, PositiveInt


class _ScheduleConfig(BaseModel):
    stages_per_rank_fixed: ClassVar[int | None] = None  # schedules with a structurally fixed number of stages per rank
    needs_backward: ClassVar[bool] = True

    def stages_per_rank(self) -> int:
        fixed = type(self).stages_per_rank_fixed
        return fixed if fixed is not None else int(getattr(self, "num_stages_per_rank"))


class PipelineScheduleInferenceConfig(_ScheduleConfig):
    """Forward-only execution."""

    schedule: Literal["inference"] = "inference"
    stages_per_rank_fixed = 1
    needs_backward = False


class PipelineScheduleGPipeConfig(_ScheduleConfig):
    """All forwards, then all backwards; one stage per rank."""

    schedule: Literal["gpipe"] = "gpipe"
    stages_per_rank_fixed = 1


class PipelineScheduleLoopedBFSConfig(_ScheduleConfig):
    """GPipe generalised to several (looped) stages per rank."""

    schedule: Literal["looped_bfs"] = "looped_bfs"
    num_stages_per_rank: PositiveInt


class PipelineSchedule1F1BConfig(_ScheduleConfig):
    """(Interleaved) 1F1B; ``zero_bubble`` splits the backward into dI and dW (ZB1P)."""

    schedule: Literal["1f1b"] = "1f1b"
    num_stages_per_rank: PositiveInt
    zero_bubble: bool


class PipelineScheduleZeroBubbleVConfig(_ScheduleConfig):
    """Zero-bubble on the V topology (exactly two stages per rank)."""

    schedule: Literal["zero_bubble_v"] = "zero_bubble_v"
    stages_per_rank_fixed = 2


class PipelineScheduleDualPipeVConfig(_ScheduleConfig):
    """DualPipeV: V topology with paired forward/backward slots."""

    schedule: Literal["dual_pipe_v"] = "dual_pipe_v"
    stages_per_rank_fixed = 2


AnyPipelineScheduleConfig = Annotated[
    PipelineSchedule1F1BConfig
    | PipelineScheduleDualPipeVConfig
    | PipelineScheduleGPipeConfig
    | PipelineScheduleInferenceConfig
    | PipelineScheduleLoopedBFSConfig
    | PipelineScheduleZeroBubbleVConfig,
    Field(discriminator="schedule"),
]
