"""User-facing pipelining API (reference ``d9d/pipelining/api``)."""

from .module import ModuleSupportsPipelining, PipelineStageInfo, distribute_layers_for_pipeline_stage
from .schedule import PipelineSchedule
from .sharding import PipelineShardingSpec
from .types import PipelineLossFn, PipelineResultFn

__all__ = [
    "ModuleSupportsPipelining",
    "PipelineLossFn",
    "PipelineResultFn",
    "PipelineSchedule",
    "PipelineShardingSpec",
    "PipelineStageInfo",
    "distribute_layers_for_pipeline_stage",
]
