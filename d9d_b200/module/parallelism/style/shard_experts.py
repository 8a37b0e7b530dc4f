from __future__ import annotations

from torch import nn
from torch.distributed import DeviceMesh
from torch.distributed.tensor import Replicate, Shard, distribute_tensor
from torch.distributed.tensor.parallel import ParallelStyle

from d9d_b200.module.block.moe import GroupedLinear, MoELayer


class ShardMoESparseExpertsParallel(ParallelStyle):
    """Expert parallelism for a :class:`MoELayer`: every ``GroupedLinear.weight [E, in, out]`` becomes a DTensor
    ``Shard(0)`` on ``shard_dim_name`` and ``Replicate`` on the other mesh dims; if that dim is larger than one the
    layer switches to the distributed dispatch/combine handler.

    Parity: reference ``d9d/module/parallelism/style/shard_experts.py:14-56``.
    """

    def __init__(self, shard_dim_name: str):
        self._shard_dim_name = shard_dim_name

    def _apply(self, module: nn.Module, device_mesh: DeviceMesh) -> nn.Module:
        if not isinstance(module, MoELayer):
            raise TypeError("This plan should be applied only on MoELayer")
        names = device_mesh.mesh_dim_names
        if names is None:
            raise ValueError("This plan should be applied only on named DeviceMeshes")
        placements = [Shard(0) if n == self._shard_dim_name else Replicate() for n in names]
        if device_mesh[self._shard_dim_name].size() > 1:
            module.enable_distributed_communicator(device_mesh.get_group(self._shard_dim_name))
        for sub in module.modules():
            if isinstance(sub, GroupedLinear):
                sub.weight = nn.Parameter(
                    distribute_tensor(sub.weight.data, device_mesh, placements, src_data_rank=None),
                    requires_grad=sub.weight.requires_grad,
                )
        return module
