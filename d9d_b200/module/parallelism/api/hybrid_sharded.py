from typing import Any

from torch import nn
from torch.distributed import DeviceMesh

from .fully_sharded import parallelize_fsdp
from .replicate_parallel import parallelize_replicate


def parallelize_hsdp(module: nn.Module, mesh: DeviceMesh, shard_dim: str = "dp_cp_shard", *fsdp_args: Any,
                     skip_distributed: bool = False, **fsdp_kwargs: Any) -> None:
    """Hybrid sharding: replicate over every mesh dim other than ``shard_dim`` (size > 1), FSDP over ``shard_dim``
    (size > 1).  ``skip_distributed``: parameters another style already distributed (tensor-parallel linears) are not
    replicated again - FSDP still shards them along ``shard_dim``.
    Parity: reference ``d9d/module/parallelism/api/hybrid_sharded.py:10-45``."""
    names = mesh.mesh_dim_names
    if names is None:
        raise ValueError("Cannot use with unnamed device meshes")
    replicate_dims = tuple(n for n in names if n != shard_dim and mesh[n].size() > 1)
    if replicate_dims:
        parallelize_replicate(module, mesh[replicate_dims], skip_distributed=skip_distributed)
    if mesh[shard_dim].size() != 1:
        parallelize_fsdp(module, mesh[shard_dim], *fsdp_args, **fsdp_kwargs)
