from __future__ import annotations

from collections.abc import Iterator
from pathlib import Path

import torch
from torch import nn
from torch.distributed import DeviceMesh
from torch.distributed.tensor import DTensor

from d9d_b200.model_state.mapper import ModelStateMapper
from d9d_b200.model_state.mapper.compose import ModelStateMapperParallel, ModelStateMapperSequential
from d9d_b200.model_state.mapper.leaf import ModelStateMapperGatherFullTensor, ModelStateMapperIdentity

from .writer import write_model_state_local, write_model_state_pipeline_parallel


def _gather_stage(models: list[nn.Module], mapper: ModelStateMapper) -> ModelStateMapper:
    """First stage turning DTensor states into full tensors before the user's mapper sees them."""
    current: dict[str, torch.Tensor] = {}
    for m in models:
        current.update(m.state_dict())
    leaves: list[ModelStateMapper] = []
    for name in sorted(mapper.all_inputs()):
        leaves.append(ModelStateMapperGatherFullTensor(name) if isinstance(current[name], DTensor) else ModelStateMapperIdentity(name))
    return ModelStateMapperSequential([ModelStateMapperParallel(leaves), mapper])


def _states(models: list[nn.Module]) -> Iterator[tuple[str, torch.Tensor]]:
    for m in models:
        yield from m.state_dict().items()


def save_model_state(dest_dir: Path, mapper: ModelStateMapper, model: nn.Module, shard_size_gb: float = 4.0,
                     show_progress: bool = True) -> None:
    """Save ``model`` from a single process.  Only keys consumed by ``mapper`` are saved."""
    write_model_state_local(dest_dir=dest_dir, mapper=_gather_stage([model], mapper), state_generator=_states([model]),
                            shard_size_gb=shard_size_gb, show_progress=show_progress)


def save_model_state_pipeline_parallel(dest_dir: Path, mapper: ModelStateMapper, device_mesh: DeviceMesh, pipeline_dim_name: str,
                                       models: list[nn.Module], shard_size_gb: float = 4.0, show_progress: bool = True,
                                       position: int | None = None) -> None:
    """Save the stages held by this rank in an N-D (pp x ...) job: DTensors are gathered, one writer per pp rank."""
    write_model_state_pipeline_parallel(dest_dir=dest_dir, mapper=_gather_stage(models, mapper), state_generator=_states(models),
                                        device_mesh=device_mesh, pipeline_dim_name=pipeline_dim_name,
                                        shard_size_gb=shard_size_gb, show_progress=show_progress, position=position)
