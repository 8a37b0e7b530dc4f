from __future__ import annotations

import dataclasses
from collections.abc import Iterable

import torch
from torch import nn
from torch.distributed import DeviceMesh
from torch.distributed.tensor import DTensor

from d9d_b200.internals.grad_sync.placement import is_shard_placement


@dataclasses.dataclass(kw_only=True, frozen=True)
class GradNormGroup:
    """Parameters whose local norms can be combined with one collective: same shard meshes, device, grad dtype."""

    shard_meshes: tuple[DeviceMesh, ...] | None
    device: torch.device
    grad_dtype: torch.dtype | None


ParametersForNorm = dict[GradNormGroup, list[nn.Parameter]]


def _shard_meshes(param: nn.Parameter) -> tuple[DeviceMesh, ...] | None:
    data = param.data
    if not isinstance(data, DTensor):
        return None
    names = data.device_mesh.mesh_dim_names
    if names is None:
        raise ValueError("Only named meshes are supported.")
    sharded = tuple(data.device_mesh[names[i]] for i, pl in enumerate(data.placements) if is_shard_placement(pl))
    return sharded or None


def group_parameters_for_norm(parameters: Iterable[nn.Parameter]) -> ParametersForNorm:
    """Group trainable parameters by the meshes they are sharded on; sharded groups first (their collectives are
    launched early and overlap the local norms of the rest)."""
    groups: ParametersForNorm = {}
    for p in parameters:
        if not p.requires_grad:
            continue
        key = GradNormGroup(shard_meshes=_shard_meshes(p), device=p.device, grad_dtype=p.grad_dtype)
        groups.setdefault(key, []).append(p)
    return dict(sorted(groups.items(), key=lambda kv: kv[0].shard_meshes is None))
