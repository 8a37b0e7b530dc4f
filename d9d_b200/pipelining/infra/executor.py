from __future__ import annotations

from typing import Any

import torch
import torch.distributed as dist
from torch import nn
from torch.autograd.profiler import record_function

from d9d_b200.core.dist_context import REGULAR_DOMAIN, DistributedContext
from d9d_b200.core.sharding import ShardingSpec, shard_spec_on_dim, shard_tree
from d9d_b200.pipelining.api import PipelineLossFn, PipelineResultFn, PipelineSchedule, PipelineShardingSpec

from .action import Action, ActionKind, Program, flatten
from .stage import PipelineStage


class PipelineLossHandler:
    """Computes the loss when the last stage finishes a microbatch and hands it to the matching backward."""

    def __init__(self, loss_fn: PipelineLossFn):
        self._loss_fn = loss_fn
        self._losses: dict[int, torch.Tensor] = {}

    def trigger(self, forward_result: dict[str, torch.Tensor], microbatch_index: int) -> None:
        self._losses[microbatch_index] = self._loss_fn(forward_result, microbatch_index)

    def acquire_loss(self, microbatch_index: int) -> torch.Tensor:
        if microbatch_index not in self._losses:
            raise ValueError(f"loss of microbatch {microbatch_index} was not computed")
        return self._losses.pop(microbatch_index)


class PipelineResultHandler:
    def __init__(self, callback_fn: PipelineResultFn):
        self._fn = callback_fn

    def trigger(self, forward_result: dict[str, torch.Tensor], microbatch_index: int) -> None:
        self._fn(forward_result, microbatch_index)


class _Transport:
    """Issues batched isend/irecv and remembers the outstanding work per (direction, stage, microbatch)."""

    def __init__(self, stages: dict[int, PipelineStage]):
        self._stages = stages
        self._recv: dict[tuple[str, int, int], list[dist.Work]] = {}
        self._send: list[dist.Work] = []

    @staticmethod
    def _launch(ops: list[dist.P2POp]) -> list[dist.Work]:
        return dist.batch_isend_irecv(ops) if ops else []

    def post(self, act: Action) -> None:
        stage = self._stages[act.stage]
        if act.kind == ActionKind.RECV_F:
            self._recv[("f", act.stage, act.microbatch)] = self._launch(stage.get_fwd_recv_ops(act.microbatch))
        elif act.kind == ActionKind.RECV_B:
            self._recv[("b", act.stage, act.microbatch)] = self._launch(stage.get_bwd_recv_ops(act.microbatch))
        elif act.kind == ActionKind.SEND_F:
            self._send += self._launch(stage.get_fwd_send_ops(act.microbatch))
        elif act.kind == ActionKind.SEND_B:
            self._send += self._launch(stage.get_bwd_send_ops(act.microbatch))

    def wait_recv(self, direction: str, stage: int, microbatch: int) -> None:
        for w in self._recv.pop((direction, stage, microbatch), []):
            w.wait()

    def wait_send_all(self) -> None:
        for w in self._send:
            w.wait()
        self._send = []
        if self._recv:
            raise RuntimeError(f"pipeline step finished with unconsumed receives: {sorted(self._recv)}")


class PipelineScheduleExecutor(PipelineSchedule):
    """Interprets this rank's action list (reference ``component/runtime/executor.py:19-112``)."""

    def __init__(self, dist_context: DistributedContext, stages: list[PipelineStage], num_microbatches: int,
                 callback: PipelineLossFn | PipelineResultFn, program: Program):
        self._ctx = dist_context
        self._stages = {s.info.current_stage: s for s in stages}
        self._num_microbatches = num_microbatches
        self._program = program
        self._has_backward = any(a.has_backward_work for actions in program.values() for a in actions)
        self._transport = _Transport(self._stages)
        self._callback: PipelineLossHandler | PipelineResultHandler = (
            PipelineLossHandler(callback) if self._has_backward else PipelineResultHandler(callback)
        )
        self._data_spec: ShardingSpec | None = None
        self._kwargs_spec: ShardingSpec | None = None
        self._buffers_key: Any = None
        mesh = dist_context.mesh_for(REGULAR_DOMAIN)
        self._pp_rank = mesh.get_local_rank("pp")

    def configure_buffers(self, inputs: dict[str, torch.Tensor], kwargs: dict[str, Any], sharding_spec: PipelineShardingSpec | None) -> None:
        self._data_spec = sharding_spec.input_data if sharding_spec and sharding_spec.input_data is not None else shard_spec_on_dim(inputs, dim=0)
        self._kwargs_spec = sharding_spec.input_kwargs if sharding_spec and sharding_spec.input_kwargs is not None else shard_spec_on_dim(kwargs, dim=0)
        # boundary metadata only depends on shapes/dtypes: re-plan only when they change (the reference re-plans
        # and re-allocates on every call)
        key = tuple(sorted((k, tuple(v.shape), v.dtype) for k, v in inputs.items()))
        if key != self._buffers_key:
            for stage in self._stages.values():
                stage.configure_buffers(num_microbatches=self._num_microbatches, pipeline_inputs=inputs, has_backward=self._has_backward)
            self._buffers_key = key

    def _run_compute(self, act: Action, inputs_mb, kwargs_mb) -> None:
        stage = self._stages[act.stage]
        s, m = act.stage, act.microbatch
        if act.kind == ActionKind.FORWARD:
            if not stage.info.is_current_stage_first and (s - 1) not in self._stages:
                self._transport.wait_recv("f", s, m)
            stage.forward_one_chunk(microbatch_index=m, pipeline_inputs=inputs_mb[m], pipeline_kwargs=kwargs_mb[m])
            result = stage.get_local_fwd_output(m)
            if stage.info.is_current_stage_last:
                self._callback.trigger(result, m)
            elif (s + 1) in self._stages:
                self._stages[s + 1].set_local_fwd_input(inputs=result, microbatch_index=m)
        elif act.kind in (ActionKind.BACKWARD_FULL, ActionKind.BACKWARD_INPUT):
            if not stage.info.is_current_stage_last and (s + 1) not in self._stages:
                self._transport.wait_recv("b", s, m)
            loss = None
            if stage.info.is_current_stage_last and isinstance(self._callback, PipelineLossHandler):
                loss = self._callback.acquire_loss(m)
            stage.backward_one_chunk(microbatch_index=m, full_backward=act.kind == ActionKind.BACKWARD_FULL, loss=loss)
            if not stage.info.is_current_stage_first and (s - 1) in self._stages:
                self._stages[s - 1].set_local_bwd_input(microbatch_index=m, inputs=stage.pop_local_bwd_output(m))
        elif act.kind == ActionKind.BACKWARD_WEIGHT:
            stage.backward_weight_one_chunk(microbatch_index=m)

    def step(self, inputs: dict[str, torch.Tensor], kwargs: dict[str, Any]) -> None:
        if self._data_spec is None or self._kwargs_spec is None:
            raise ValueError("Please configure sharding specs first")
        for stage in self._stages.values():
            stage.reset()
        inputs_mb = shard_tree(inputs, self._data_spec, num_shards=self._num_microbatches, enforce_even_split=True)
        kwargs_mb = shard_tree(kwargs, self._kwargs_spec, num_shards=self._num_microbatches, enforce_even_split=True)
        for action in self._program[self._pp_rank]:
            with record_function(str(action)):
                for act in flatten(action):
                    if act.is_compute:
                        self._run_compute(act, inputs_mb, kwargs_mb)
                    else:
                        self._transport.post(act)
        self._transport.wait_send_all()
        for stage in self._stages.values():
            stage.finish_step()
            stage.assert_drained()


class OfflinePipelineExecutor(PipelineSchedule):
    """No pipeline at all: one stage, one microbatch, forward -> callback -> backward
    (reference ``component/runtime/offline.py:47-54``)."""

    def __init__(self, model: nn.Module, callback: PipelineLossFn | PipelineResultFn, do_backward: bool):
        self._model = model
        self._callback = callback
        self._do_backward = do_backward

    def configure_buffers(self, inputs, kwargs, sharding_spec) -> None:
        return None

    def step(self, inputs: dict[str, torch.Tensor], kwargs: dict[str, Any]) -> None:
        result = self._model(**inputs, **kwargs)
        result = {k: v for k, v in result.items() if v is not None}
        out = self._callback(result, 0)
        if self._do_backward:
            out.backward()
