from __future__ import annotations

from pathlib import Path

import torch
from torch import nn
from torch.distributed.tensor import DTensor

from d9d_b200.model_state.mapper import ModelStateMapper
from d9d_b200.model_state.mapper.compose import ModelStateMapperParallel, ModelStateMapperSequential
from d9d_b200.model_state.mapper.leaf import ModelStateMapperDistribute, ModelStateMapperIdentity

from .reader import read_model_state


def _placement_stage(model: nn.Module, mapper: ModelStateMapper) -> ModelStateMapper:
    """Second stage turning loaded full tensors into DTensors matching the model's parameters (no communication)."""
    current = model.state_dict()
    leaves: list[ModelStateMapper] = []
    for name in sorted(mapper.all_outputs()):
        target = current[name]
        if isinstance(target, DTensor):
            leaves.append(ModelStateMapperDistribute(name=name, device_mesh=target.device_mesh, placements=target.placements))
        else:
            leaves.append(ModelStateMapperIdentity(name))
    return ModelStateMapperSequential([mapper, ModelStateMapperParallel(leaves)])


def load_model_state(src_dir: Path, mapper: ModelStateMapper, device: str, model: nn.Module, show_progress: bool = True,
                     position: int | None = None) -> None:
    """Stream a checkpoint into ``model``: map names/layouts with ``mapper``, shard to the parameters' DTensor
    placements, inject tensor by tensor with ``load_state_dict(strict=False)``.  Only keys the mapper produces load."""
    full_mapper = _placement_stage(model, mapper)
    for name, value in read_model_state(src_dir=src_dir, mapper=full_mapper, device=device, show_progress=show_progress, position=position):
        model.load_state_dict({name: value}, strict=False)
