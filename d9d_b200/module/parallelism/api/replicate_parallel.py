from torch import nn
from torch.distributed import DeviceMesh
from torch.distributed.tensor import Replicate
from torch.distributed.tensor.parallel import parallelize_module

from d9d_b200.module.parallelism.style import ToLocalParallel


def parallelize_replicate(module: nn.Module, mesh: DeviceMesh, skip_distributed: bool = False) -> None:
    """Data-parallel replication expressed with DTensors: parameters become ``DTensor(Replicate x ndim)``, forward runs
    on local tensors, gradients come back as ``Replicate`` DTensors that the ``GradientSynchronizer`` SUM-reduces.

    ``skip_distributed=True`` leaves parameters that are already ``DTensor`` s alone (used after tensor-parallel styles
    claimed a module's linears: the remaining parameters - norms, gates - are replicated).

    Parity: reference ``d9d/module/parallelism/api/replicate_parallel.py:9-37``.
    """
    placement = tuple(Replicate() for _ in range(mesh.ndim))
    parallelize_module(module, mesh, ToLocalParallel(param_placement=placement, grad_placement=placement,
                                                    skip_distributed=skip_distributed))
