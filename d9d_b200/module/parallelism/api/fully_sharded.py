from typing import Any

from torch import nn
from torch.distributed import DeviceMesh
from torch.distributed.fsdp import FSDPModule, fully_shard


def parallelize_fsdp(module: nn.Module, mesh: DeviceMesh, *args: Any, **kwargs: Any) -> None:
    """FSDP2 ``fully_shard`` over a 1-D mesh with SUM gradient reduction (no averaging, no internal all-reduce):
    normalisation by the global loss weight and replica reduction are done by the training loop.

    Parity: reference ``d9d/module/parallelism/api/fully_sharded.py:8-39``.
    """
    if mesh.ndim != 1:
        raise ValueError("FSDP mesh should contain exactly one dimension - for HSDP, please apply parallelize_replicate(...) first!")
    fully_shard(module, *args, mesh=mesh, **kwargs)
    if not isinstance(module, FSDPModule):
        raise RuntimeError("Torch FSDP did not convert the module into FSDPModule")
    module.set_force_sum_reduction_for_comms(enable=True)
    module.set_gradient_divide_factor(1.0)
    module.set_requires_all_reduce(False)
