from torch.distributed import DeviceMesh
from torch.distributed.tensor import Replicate
from torch.distributed.tensor.parallel import parallelize_module

from d9d_b200.module.block.moe import MoELayer
from d9d_b200.module.parallelism.style import ShardMoESparseExpertsParallel, ToLocalParallel


def parallelize_expert_parallel(module: MoELayer, mesh_experts: DeviceMesh, expert_shard_dim: str = "ep_shard") -> None:
    """Shard the grouped experts over ``expert_shard_dim`` and replicate the router (and shared expert) over the
    whole expert mesh.  Parity: reference ``d9d/module/parallelism/api/expert_parallel.py:9-47``."""
    parallelize_module(module, mesh_experts, ShardMoESparseExpertsParallel(shard_dim_name=expert_shard_dim))
    replicate = tuple(Replicate() for _ in range(mesh_experts.ndim))
    parallelize_module(module.router, mesh_experts, ToLocalParallel(param_placement=replicate, grad_placement=replicate))
    if module.shared_expert is not None:
        parallelize_module(module.shared_expert, mesh_experts, ToLocalParallel(param_placement=replicate, grad_placement=replicate))
