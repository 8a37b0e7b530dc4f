from __future__ import annotations

import dataclasses
from typing import Protocol, runtime_checkable

import torch


@dataclasses.dataclass
class PipelineStageInfo:
    """Position of a model chunk in the (virtual) pipeline."""

    current_stage: int
    num_stages: int

    @property
    def is_current_stage_first(self) -> bool:
        return self.current_stage == 0

    @property
    def is_current_stage_last(self) -> bool:
        return self.current_stage == self.num_stages - 1


def layers_per_stage(num_layers: int, num_virtual_layers_pre: int, num_virtual_layers_post: int, num_stages: int) -> list[int]:
    """Real layer count of every stage.

    ``pre`` / ``post`` virtual layers stand for the cost of embeddings / final norm + head: the virtual total is
    split as evenly as possible (remainder to the earliest stages) and the virtual ones are then subtracted from
    the first / last stage.
    """
    virtual_total = num_layers + num_virtual_layers_pre + num_virtual_layers_post
    base, extra = divmod(virtual_total, num_stages)
    counts = []
    for s in range(num_stages):
        n = base + (1 if s < extra else 0)
        if s == 0:
            n -= num_virtual_layers_pre
        if s == num_stages - 1:
            n -= num_virtual_layers_post
        if n <= 0:
            raise ValueError(
                f"Tried to distribute layers, but got {n} on stage {s}. Perhaps the pipeline is too long for this model?"
            )
        counts.append(n)
    return counts


def distribute_layers_for_pipeline_stage(num_layers: int, num_virtual_layers_pre: int, num_virtual_layers_post: int,
                                         stage: PipelineStageInfo) -> tuple[int, int]:
    """``[start, end)`` global layer indices owned by ``stage`` (reference ``pipelining/api/module.py:43-105``)."""
    counts = layers_per_stage(num_layers, num_virtual_layers_pre, num_virtual_layers_post, stage.num_stages)
    start = sum(counts[: stage.current_stage])
    return start, start + counts[stage.current_stage]


@runtime_checkable
class ModuleSupportsPipelining(Protocol):
    """Modules that can describe their stage-local input/output tensors (as meta/empty tensors of *microbatch*
    shape) from the global pipeline inputs, so p2p buffers can be planned without running a forward."""

    def infer_stage_inputs_from_pipeline_inputs(self, inputs: dict[str, torch.Tensor], n_microbatches: int) -> dict[str, torch.Tensor]: ...

    def infer_stage_outputs_from_pipeline_inputs(self, inputs: dict[str, torch.Tensor], n_microbatches: int) -> dict[str, torch.Tensor]: ...
