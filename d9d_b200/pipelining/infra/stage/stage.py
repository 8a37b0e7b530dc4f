from __future__ import annotations

import dataclasses
import os
from collections.abc import Mapping
from typing import Any

import torch
import torch.distributed as dist
from torch import nn

from d9d_b200.pipelining.api import ModuleSupportsPipelining, PipelineStageInfo

from .p2p import BoundaryChannel
from .split_backward import DeferredWeightBackward, backward_full, backward_input, backward_weight


@dataclasses.dataclass
class _ForwardRecord:
    inputs: dict[str, torch.Tensor]
    outputs: dict[str, torch.Tensor]


@dataclasses.dataclass
class _DeferredFull:
    outputs: list[torch.Tensor]
    output_grads: list[torch.Tensor] | None
    inputs: list[torch.Tensor]


class _AlreadyDone:
    """Placeholder for a weight pass whose work was done together with the input pass."""


class _FsdpSplitBackward:
    """Split backward (input pass now, weight pass later) over FSDP2-managed stages.

    FSDP frees the unsharded parameters and reduce-scatters their gradients from hooks that fire when a backward pass
    reaches the module inputs - i.e. during the *input* pass, before any weight gradient exists, after which the deferred
    weight pass would read freed storage.  While a stage has split backward work in flight this guard therefore switches
    FSDP into its gradient-accumulation mode (no reshard after forward / backward, no gradient reduction, "not the last
    backward"): parameters stay unsharded, weight gradients accumulate on the unsharded parameters.  ``release()`` - called
    by the executor at the end of the pipeline step - restores the flags and runs FSDP's post-backward once: one
    reduce-scatter per step, parameters resharded.  The same protocol as ``torch.distributed.pipelining`` uses
    (``stage.py: backward_maybe_with_nosync / perform_reduce_grad``); it relies on FSDP2's private state objects, so any
    failure to find them falls back to running the whole backward in the input slot.
    """

    def __init__(self, module: nn.Module):
        from torch.distributed.fsdp import FSDPModule

        self._modules = [m for m in module.modules() if isinstance(m, FSDPModule)]
        self._saved: list[tuple] | None = None
        self.supported = bool(self._modules) and os.environ.get("D9D_PP_FSDP_SPLIT", "1") != "0"
        if self.supported:
            try:
                for m in self._modules:
                    state = m._get_fsdp_state()  # noqa: SLF001
                    _ = state._state_ctx, state._fsdp_param_group, state._root_post_backward_final_callback  # noqa: SLF001
            except AttributeError:
                self.supported = False

    def __bool__(self) -> bool:
        return bool(self._modules)

    @property
    def engaged(self) -> bool:
        return self._saved is not None

    def engage(self) -> None:
        if self._saved is not None:
            return
        saved = []
        for m in self._modules:
            state = m._get_fsdp_state()  # noqa: SLF001
            group = state._fsdp_param_group  # noqa: SLF001
            saved.append((state, state._auto_reshard_after_forward,  # noqa: SLF001
                          None if group is None else (group.post_forward_mesh_info, group.reshard_after_backward, group.reduce_grads,
                                                      group.all_reduce_grads)))
            m.set_is_last_backward(False)
            m.set_reshard_after_backward(False, recurse=False)
            m.set_requires_gradient_sync(False, recurse=False)
            m.set_reshard_after_forward(False, recurse=False)
        self._saved = saved

    def unshard(self) -> None:
        """All-gather (a no-op when already done) and register the unsharded parameters on the modules: the autograd graph of a
        forward pass hangs off *those* tensors, and the split-backward analysis looks parameters up through the modules."""
        for m in self._modules:
            m.unshard()

    def release(self) -> None:
        if self._saved is None:
            return
        saved, self._saved = self._saved, None
        for m, (state, auto, group_flags) in zip(self._modules, saved, strict=True):
            state._auto_reshard_after_forward = auto  # noqa: SLF001
            group = state._fsdp_param_group  # noqa: SLF001
            if group is not None and group_flags is not None:
                group.post_forward_mesh_info, group.reshard_after_backward, group.reduce_grads, group.all_reduce_grads = group_flags
            m.set_is_last_backward(True)
        roots = []
        for state, _auto, _flags in saved:
            if state._fsdp_param_group is not None:  # noqa: SLF001
                state._fsdp_param_group.post_backward()  # noqa: SLF001  reduce-scatter of the accumulated gradients, reshard
            if getattr(state._state_ctx, "all_states", None) and state._state_ctx.all_states[0] is state:  # noqa: SLF001
                roots.append(state)
        for state in roots or [saved[0][0]]:
            state._root_post_backward_final_callback()  # noqa: SLF001  joins the reduction streams, resets FSDP's step state


class PipelineStage:
    """One model chunk of the pipeline: runs forward / (split) backward per microbatch and owns its boundaries.

    Parity: reference ``d9d/pipelining/infra/stage/stage.py:22-293`` + ``computations.py`` + ``communications.py``.
    """

    def __init__(self, info: PipelineStageInfo, module: nn.Module, group: dist.ProcessGroup | None,
                 stage_to_host_topology: dict[int, int]):
        self._info = info
        self._module = module
        self._group = group
        self._topology = stage_to_host_topology
        self._has_backward = False
        self._fwd_in: BoundaryChannel | None = None  # activations from the previous stage
        self._bwd_in: BoundaryChannel | None = None  # output gradients from the next stage
        self._forward: dict[int, _ForwardRecord] = {}
        self._input_grads: dict[int, dict[str, torch.Tensor | None]] = {}
        self._deferred: dict[int, DeferredWeightBackward | _DeferredFull | _AlreadyDone] = {}
        self._can_split_backward: bool | None = None
        self._fsdp: _FsdpSplitBackward | None = None

    @property
    def info(self) -> PipelineStageInfo:
        return self._info

    @property
    def module(self) -> nn.Module:
        return self._module

    def _peer(self, stage: int | None) -> int | None:
        if stage is None or self._group is None:
            return None
        me = self._topology[self._info.current_stage]
        other = self._topology[stage]
        return None if other == me else other

    def configure_buffers(self, num_microbatches: int, has_backward: bool, pipeline_inputs: dict[str, torch.Tensor]) -> None:
        """(Re)plan boundary metadata for inputs of this shape."""
        if not isinstance(self._module, ModuleSupportsPipelining):
            raise TypeError("Module does not implement ModuleSupportsPipelining protocol")
        self._has_backward = has_backward
        prev_stage = None if self._info.is_current_stage_first else self._info.current_stage - 1
        next_stage = None if self._info.is_current_stage_last else self._info.current_stage + 1
        device = next((p.device for p in self._module.parameters()), torch.device("cpu"))
        with torch.device("meta"):
            meta_inputs = {k: torch.empty(v.shape, dtype=v.dtype, device="meta") for k, v in pipeline_inputs.items()}
            in_meta = self._module.infer_stage_inputs_from_pipeline_inputs(inputs=meta_inputs, n_microbatches=num_microbatches)
            out_meta = self._module.infer_stage_outputs_from_pipeline_inputs(inputs=meta_inputs, n_microbatches=num_microbatches)
        self._fwd_in = BoundaryChannel("fwd", self._info.current_stage, self._peer(prev_stage), self._group,
                                       in_meta if prev_stage is not None else {}, device, requires_grad=has_backward)
        self._bwd_in = BoundaryChannel("bwd", self._info.current_stage, self._peer(next_stage), self._group,
                                       {k: v for k, v in out_meta.items() if v.is_floating_point()} if next_stage is not None else {},
                                       device, requires_grad=False) if has_backward else None
        self._send_fwd = BoundaryChannel("fwd-out", self._info.current_stage, self._peer(next_stage), self._group, {}, device, False)
        self._send_bwd = BoundaryChannel("bwd-out", self._info.current_stage, self._peer(prev_stage), self._group, {}, device, False)

    # ------------------------------------------------------------------ p2p op factories
    def get_fwd_recv_ops(self, microbatch: int) -> list[dist.P2POp]:
        assert self._fwd_in is not None, "You must configure stage buffers first"
        return self._fwd_in.make_recv_ops(microbatch)

    def get_fwd_send_ops(self, microbatch: int) -> list[dist.P2POp]:
        outputs = self._forward[microbatch].outputs
        return self._send_fwd.make_send_ops(outputs)

    def get_bwd_recv_ops(self, microbatch: int) -> list[dist.P2POp]:
        if not self._has_backward or self._bwd_in is None:
            return []
        return self._bwd_in.make_recv_ops(microbatch)

    def get_bwd_send_ops(self, microbatch: int) -> list[dist.P2POp]:
        if not self._has_backward:
            return []
        grads = self.pop_local_bwd_output(microbatch)
        return self._send_bwd.make_send_ops({k: v for k, v in grads.items() if v is not None})

    # ------------------------------------------------------------------ same-rank hand-off
    def set_local_fwd_input(self, inputs: dict[str, torch.Tensor], microbatch_index: int) -> None:
        assert self._fwd_in is not None, "You must configure stage buffers first"
        self._fwd_in.set_local(inputs, microbatch_index)

    def get_local_fwd_output(self, microbatch_index: int) -> dict[str, torch.Tensor]:
        return self._forward[microbatch_index].outputs

    def set_local_bwd_input(self, inputs: dict[str, torch.Tensor | None], microbatch_index: int) -> None:
        assert self._bwd_in is not None, "You must configure stage buffers first"
        self._bwd_in.set_local({k: v for k, v in inputs.items() if v is not None}, microbatch_index)

    def pop_local_bwd_output(self, microbatch_index: int) -> dict[str, torch.Tensor | None]:
        return self._input_grads.pop(microbatch_index)

    # ------------------------------------------------------------------ compute
    def forward_one_chunk(self, microbatch_index: int, pipeline_inputs: dict[str, torch.Tensor],
                          pipeline_kwargs: dict[str, Any] | None = None) -> None:
        assert self._fwd_in is not None, "You must configure stage buffers first"
        inputs = pipeline_inputs if self._info.is_current_stage_first else self._fwd_in.take(microbatch_index)
        try:
            output = self._module(**inputs, **(pipeline_kwargs or {}))
        except Exception as exc:
            raise RuntimeError(f"S{self._info.current_stage}B{microbatch_index} failed to run forward") from exc
        if not isinstance(output, Mapping):
            raise ValueError("Currently, pipelined models should output dict[str, torch.Tensor | None]")
        self._forward[microbatch_index] = _ForwardRecord(inputs=inputs, outputs={k: v for k, v in output.items() if v is not None})

    def _trainable(self) -> list[nn.Parameter]:
        return [p for p in self._module.parameters() if p.requires_grad]

    def backward_one_chunk(self, microbatch_index: int, loss: torch.Tensor | None = None, full_backward: bool = True) -> None:
        if not self._has_backward:
            raise ValueError("stage was configured without backward")
        if microbatch_index in self._deferred or microbatch_index in self._input_grads:
            raise ValueError(f"S{self._info.current_stage}B{microbatch_index} double backward")
        record = self._forward.pop(microbatch_index)
        if self._info.is_current_stage_last:
            if loss is None:
                raise ValueError("Cannot perform backward on last stage without loss specified")
            outs, out_grads = [loss], None
        else:
            assert self._bwd_in is not None
            grads_in = self._bwd_in.take(microbatch_index)
            keys = [k for k in sorted(record.outputs) if k in grads_in and record.outputs[k].requires_grad]
            outs = [record.outputs[k] for k in keys]
            out_grads = [grads_in[k] for k in keys]
        in_keys = sorted(record.inputs)
        in_list = [record.inputs[k] for k in in_keys]

        if self._can_split_backward is None:
            self._fsdp = _FsdpSplitBackward(self._module)
            self._can_split_backward = (not self._fsdp) or self._fsdp.supported
        if not full_backward and self._fsdp:
            if self._can_split_backward:
                self._fsdp.engage()  # parameters stay unsharded, gradients unreduced, until the end of this pipeline step
                self._fsdp.unshard()
        if not full_backward and not self._can_split_backward:
            # FSDP-sharded stage without the private hooks this needs: the input slot of a zero-bubble schedule runs the whole
            # backward (correct, merely without the bubble-filling benefit on this stage); the weight slot becomes a no-op
            if os.environ.get("D9D_PP_FSDP_SPLIT") == "require":
                raise RuntimeError("split backward over FSDP parameters was required but FSDP's state objects were not found")
            grads = backward_full(outs, out_grads, in_list)
            self._deferred[microbatch_index] = _AlreadyDone()
            if not self._info.is_current_stage_first:
                self._input_grads[microbatch_index] = dict(zip(in_keys, grads, strict=True))
        elif full_backward:
            grads = backward_full(outs, out_grads, in_list)
            if not self._info.is_current_stage_first:
                self._input_grads[microbatch_index] = dict(zip(in_keys, grads, strict=True))
        elif self._info.is_current_stage_first or not any(t.requires_grad for t in in_list):
            # nothing to send upstream: the whole backward is deferred to the W slot
            self._deferred[microbatch_index] = _DeferredFull(outs, out_grads, in_list)
            if not self._info.is_current_stage_first:
                self._input_grads[microbatch_index] = dict.fromkeys(in_keys)
        else:
            grads, deferred = backward_input(outs, out_grads, in_list, self._trainable())
            self._deferred[microbatch_index] = deferred
            self._input_grads[microbatch_index] = dict(zip(in_keys, grads, strict=True))
        if self._info.is_current_stage_last and not self._info.is_current_stage_first:
            for t in record.outputs.values():
                if not t._is_view():  # noqa: SLF001  free the graph hanging off outputs nobody will backprop through again
                    t.detach_()

    def backward_weight_one_chunk(self, microbatch_index: int) -> None:
        if microbatch_index not in self._deferred:
            raise ValueError(f"S{self._info.current_stage}W{microbatch_index} - weight backward with no input backward before")
        deferred = self._deferred.pop(microbatch_index)
        if isinstance(deferred, _AlreadyDone):
            return
        if isinstance(deferred, _DeferredFull):
            backward_full(deferred.outputs, deferred.output_grads, deferred.inputs)
        else:
            backward_weight(deferred)

    # ------------------------------------------------------------------ housekeeping
    def finish_step(self) -> None:
        """End of one pipeline step: gradients of a split backward over FSDP parameters are reduced, parameters resharded."""
        if self._fsdp is not None and self._fsdp.engaged:
            self._fsdp.release()

    def reset(self) -> None:
        if self._fsdp is not None and self._fsdp.engaged:  # a step that raised half-way: leave FSDP in a sane state
            self._fsdp.release()
        for ch in (self._fwd_in, self._bwd_in):
            if ch is not None:
                ch.reset()

    def assert_drained(self) -> None:
        """No forward / backward caches may survive a step (reference test: 'no dangling caches')."""
        if not self._has_backward:
            self._forward.clear()  # forward-only schedules keep outputs just long enough to send them
        leftovers = {"forward": len(self._forward), "input_grads": len(self._input_grads), "deferred": len(self._deferred)}
        if any(leftovers.values()):
            raise RuntimeError(f"stage {self._info.current_stage} has dangling state after the step: {leftovers}")
