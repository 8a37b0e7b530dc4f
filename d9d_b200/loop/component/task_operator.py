from __future__ import annotations

import dataclasses
from typing import Any

import torch

from d9d_b200.core.dist_context import DistributedContext
from d9d_b200.internals.pipeline_state import PipelineStateHandler
from d9d_b200.loop.control import BaseTask, BuildForwardInputsContext, InferenceTask, TrainTask, UpdateMetricsContext
from d9d_b200.metric.impl.container import ComposeMetric
from d9d_b200.pipelining.factory import PipelineScheduleInfo

from .pipeline_result_processing import STATE_LOSS, STATE_LOSS_WEIGHT


@dataclasses.dataclass(kw_only=True)
class ForwardResult:
    loss: torch.Tensor  # [n_microbatches]
    loss_weight: torch.Tensor  # [n_microbatches]


def _run(task: BaseTask, pipeline: PipelineScheduleInfo, state: PipelineStateHandler, batch: Any) -> None:
    built = task.build_forward_inputs(BuildForwardInputsContext(batch=batch, state=state.global_state()))
    pipeline.schedule.configure_buffers(inputs=built.inputs, kwargs=built.kwargs, sharding_spec=built.pipeline_sharding_spec)
    pipeline.schedule.step(inputs=built.inputs, kwargs=built.kwargs)


class TrainTaskOperator:
    """One forward+backward over a loader batch through the (possibly pipelined) schedule; updates task metrics on
    ranks holding the last stage.  The per-step state is always reset afterwards."""

    def __init__(self, dist_context: DistributedContext, task: TrainTask, pipeline: PipelineScheduleInfo,
                 pipeline_state: PipelineStateHandler, metrics: ComposeMetric):
        self._ctx, self._task, self._pipeline, self._state, self._metrics = dist_context, task, pipeline, pipeline_state, metrics

    def forward_backward(self, batch: Any) -> ForwardResult | None:
        try:
            _run(self._task, self._pipeline, self._state, batch)
            if not self._pipeline.has_last_stage:
                return None
            state = self._state.global_state()
            self._task.update_metrics(UpdateMetricsContext(state=state, metrics=self._metrics.children))
            return ForwardResult(loss=state[STATE_LOSS], loss_weight=state[STATE_LOSS_WEIGHT])
        finally:
            self._state.reset()


class InferenceTaskOperator:
    def __init__(self, dist_context: DistributedContext, task: InferenceTask, pipeline: PipelineScheduleInfo,
                 pipeline_state: PipelineStateHandler):
        self._ctx, self._task, self._pipeline, self._state = dist_context, task, pipeline, pipeline_state

    def forward(self, batch: Any) -> None:
        try:
            _run(self._task, self._pipeline, self._state, batch)
        finally:
            self._state.reset()
