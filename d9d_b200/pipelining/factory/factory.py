from __future__ import annotations

import dataclasses
from collections.abc import Callable
from typing import Any

from torch import nn

from d9d_b200.core.dist_context import REGULAR_DOMAIN, DistributedContext
from d9d_b200.pipelining.api import PipelineLossFn, PipelineResultFn, PipelineSchedule, PipelineStageInfo
from d9d_b200.pipelining.infra.executor import OfflinePipelineExecutor, PipelineScheduleExecutor
from d9d_b200.pipelining.infra.stage import PipelineStage
from d9d_b200.pipelining.infra.topology import build_stage_to_host_rank_topology, invert_stage_to_host_rank_topology

from .config import PipelineScheduleInferenceConfig
from .registry import PIPELINE_PROGRAM_REGISTRY


@dataclasses.dataclass(kw_only=True)
class PipelineScheduleInfo:
    schedule: PipelineSchedule
    has_first_stage: bool
    has_last_stage: bool


def build_schedule(dist_context: DistributedContext, n_microbatches: int, schedule_config: Any,
                   model_provider: Callable[[PipelineStageInfo], nn.Module],
                   callback: PipelineLossFn | PipelineResultFn) -> tuple[PipelineScheduleInfo, list[nn.Module]]:
    """Instantiate this rank's model stages and the schedule that drives them.

    Non-distributed context: a single stage behind an :class:`OfflinePipelineExecutor`.  Distributed: stages are
    placed by the schedule's topology, the program is composed (compute order + injected communication) and run by
    a :class:`PipelineScheduleExecutor`.  Parity: reference ``d9d/pipelining/factory/factory.py:30-131``.
    """
    if not dist_context.mesh_params.is_distributed:
        model = model_provider(PipelineStageInfo(num_stages=1, current_stage=0))
        has_backward = not isinstance(schedule_config, PipelineScheduleInferenceConfig)
        executor = OfflinePipelineExecutor(model=model, callback=callback, do_backward=has_backward)
        return PipelineScheduleInfo(schedule=executor, has_first_stage=True, has_last_stage=True), [model]

    builder = PIPELINE_PROGRAM_REGISTRY.program_for(schedule_config)
    pp_mesh = dist_context.mesh_for(REGULAR_DOMAIN)["pp"]
    pp_size = pp_mesh.size()
    num_stages = builder.num_stages_per_rank * pp_size
    stage_to_host = build_stage_to_host_rank_topology(num_stages=num_stages, pp_size=pp_size, style=builder.topology_style)
    my_stages = invert_stage_to_host_rank_topology(stage_to_host)[pp_mesh.get_local_rank()]
    group = pp_mesh.get_group() if pp_size > 1 else None

    stages, modules = [], []
    for idx in my_stages:
        info = PipelineStageInfo(num_stages=num_stages, current_stage=idx)
        module = model_provider(info)
        modules.append(module)
        stages.append(PipelineStage(info=info, module=module, group=group, stage_to_host_topology=stage_to_host))
    program = builder.compose(num_microbatches=n_microbatches, pp_size=pp_size)
    executor = PipelineScheduleExecutor(dist_context=dist_context, stages=stages, num_microbatches=n_microbatches,
                                        callback=callback, program=program)
    return PipelineScheduleInfo(schedule=executor, has_first_stage=0 in my_stages, has_last_stage=(num_stages - 1) in my_stages), modules
