import os
from typing import Any

from torch import nn
from torch.distributed import DeviceMesh
from torch.distributed.fsdp import FSDPModule, fully_shard


def parallelize_fsdp(module: nn.Module, mesh: DeviceMesh, *args: Any, peer_memory_arena_factory: Any = None, **kwargs: Any) -> None:
    """FSDP2 ``fully_shard`` over a 1-D mesh with SUM gradient reduction (no averaging, no internal all-reduce):
    normalisation by the global loss weight and replica reduction are done by the training loop.

    ``D9D_FSDP_COMM=peer`` (CUDA) replaces FSDP's NCCL all-gather / reduce-scatter by pulls over NVLink peer memory
    (``_peer_memory_fsdp.py``: opt-in, see its status note); ``peer_memory_arena_factory`` selects that path explicitly with a
    custom arena (tests emulate the peer memory on gloo).

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
    use_peer = peer_memory_arena_factory is not None or (os.environ.get("D9D_FSDP_COMM", "nccl") == "peer" and mesh.device_type == "cuda")
    if use_peer and mesh.size() > 1:
        import torch

        from ._peer_memory_fsdp import peer_memory_comms

        device = torch.device("cuda", torch.cuda.current_device()) if mesh.device_type == "cuda" else torch.device("cpu")
        all_gather, reduce_scatter = peer_memory_comms(mesh.get_group(), device, peer_memory_arena_factory)
        module.set_custom_all_gather(all_gather)
        module.set_custom_reduce_scatter(reduce_scatter)
