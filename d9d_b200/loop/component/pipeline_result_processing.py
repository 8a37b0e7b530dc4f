from __future__ import annotations

import abc
from typing import Generic, TypeVar

import torch

from d9d_b200.internals.pipeline_state import PipelineStateHandler
from d9d_b200.loop.control import ComputeLossContext, InferenceTask, ProcessOutputsContext, TrainTask

from .stepper import Stepper

STATE_LOSS = "__internal_loss"
STATE_LOSS_WEIGHT = "__internal_loss_weight"

TOutput = TypeVar("TOutput")


class PipelineOutputsProcessor(abc.ABC, Generic[TOutput]):
    @abc.abstractmethod
    def __call__(self, pipeline_outputs: dict[str, torch.Tensor], microbatch_idx: int) -> TOutput: ...


class LossComputer(PipelineOutputsProcessor[torch.Tensor]):
    """Last-stage callback: runs the task's ``compute_loss`` on the microbatch's state view, records loss and weight
    as ``[1]`` tensors (so the global view is ``[n_microbatches]``) and returns ``loss * weight`` to back-propagate."""

    def __init__(self, state: PipelineStateHandler, task: TrainTask, stepper: Stepper):
        self._state, self._task, self._stepper = state, task, stepper

    def __call__(self, pipeline_outputs: dict[str, torch.Tensor], microbatch_idx: int) -> torch.Tensor:
        view = self._state.sharded_state(shard_id=microbatch_idx)
        result = self._task.compute_loss(ComputeLossContext(pipeline_results=pipeline_outputs, state=view, stepper=self._stepper))
        weight = result.loss_weight if result.loss_weight is not None else torch.ones_like(result.loss)
        view[STATE_LOSS] = result.loss[None]
        view[STATE_LOSS_WEIGHT] = weight[None]
        return result.loss * weight


class InferenceProcessor(PipelineOutputsProcessor[None]):
    def __init__(self, state: PipelineStateHandler, task: InferenceTask):
        self._state, self._task = state, task

    def __call__(self, pipeline_outputs: dict[str, torch.Tensor], microbatch_idx: int) -> None:
        view = self._state.sharded_state(shard_id=microbatch_idx)
        self._task.process_outputs(ProcessOutputsContext(pipeline_results=pipeline_outputs, state=view))
