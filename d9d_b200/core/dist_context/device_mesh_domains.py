from __future__ import annotations

import dataclasses
from collections.abc import Callable
from typing import TYPE_CHECKING

from torch.distributed import DeviceMesh, init_device_mesh

if TYPE_CHECKING:
    from .params import DeviceMeshParameters

REGULAR_DOMAIN = "regular"
DENSE_DOMAIN = "dense"
EXPERT_DOMAIN = "expert"
BATCH_DOMAIN = "batch"
FLAT_DOMAIN = "flat"


@dataclasses.dataclass(frozen=True)
class DeviceMeshDomain:
    """A named way of folding the same row-major rank order into a mesh."""

    name: str
    dim_names: tuple[str, ...]
    shape_of: Callable[["DeviceMeshParameters"], tuple[int, ...]]

    def build_mesh(self, params: "DeviceMeshParameters", device_type: str) -> DeviceMesh:
        return init_device_mesh(device_type=device_type, mesh_shape=self.shape_of(params), mesh_dim_names=self.dim_names)


def _dp_cp(p: "DeviceMeshParameters") -> int:
    return p.data_parallel_replicate * p.data_parallel_shard * p.context_parallel_replicate * p.context_parallel_shard


ALL_DOMAIN_PROVIDERS: tuple[DeviceMeshDomain, ...] = (
    DeviceMeshDomain(
        REGULAR_DOMAIN,
        ("pp", "dp_replicate", "dp_shard", "cp_shard", "cp_replicate", "tp"),
        lambda p: (
            p.pipeline_parallel,
            p.data_parallel_replicate,
            p.data_parallel_shard,
            p.context_parallel_shard,
            p.context_parallel_replicate,
            p.tensor_parallel,
        ),
    ),
    DeviceMeshDomain(
        DENSE_DOMAIN,
        ("pp", "dp_replicate", "dp_cp_shard", "cp_replicate", "tp"),
        lambda p: (
            p.pipeline_parallel,
            p.data_parallel_replicate,
            p.data_parallel_shard * p.context_parallel_shard,
            p.context_parallel_replicate,
            p.tensor_parallel,
        ),
    ),
    DeviceMeshDomain(
        EXPERT_DOMAIN,
        ("pp", "ep_replicate", "ep_shard"),
        # NOTE: TP ranks are folded into ep_replicate so that the mesh always covers the world
        lambda p: (p.pipeline_parallel, _dp_cp(p) * p.tensor_parallel // p.expert_parallel, p.expert_parallel),
    ),
    DeviceMeshDomain(
        BATCH_DOMAIN,
        ("pp", "dp", "cp", "tp"),
        lambda p: (
            p.pipeline_parallel,
            p.data_parallel_replicate * p.data_parallel_shard,
            p.context_parallel_replicate * p.context_parallel_shard,
            p.tensor_parallel,
        ),
    ),
    DeviceMeshDomain(FLAT_DOMAIN, ("world",), lambda p: (p.pipeline_parallel * _dp_cp(p) * p.tensor_parallel,)),
)
