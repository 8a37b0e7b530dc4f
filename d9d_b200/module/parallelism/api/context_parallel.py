from __future__ import annotations

from torch import nn
from torch.distributed import DeviceMesh

from d9d_b200.kernel.context_parallel import ContextParallelLayout
from d9d_b200.module.block.attention.sdpa import ContextParallelMode, FlashSdpa


def parallelize_context_parallel(module: nn.Module, mesh_cp: DeviceMesh, mode: ContextParallelMode = ContextParallelMode.auto,
                                 layout: ContextParallelLayout = ContextParallelLayout.zigzag) -> int:
    """Make every attention kernel (:class:`FlashSdpa`) inside ``module`` attend across the ranks of the 1-D mesh
    ``mesh_cp`` (normally ``dist_context.mesh_for(BATCH_DOMAIN)["cp"]``): activations then carry ``S / cp`` tokens per
    sequence and only attention exchanges data.  Returns the number of kernels switched.

    Everything else in a decoder layer is token-wise, so no other module changes; parameters are *not* touched - they are
    replicated / FSDP-sharded over the context-parallel ranks by the dense plan (the ``cp_*`` dims are part of the dense
    mesh), which also sums their gradients over those ranks.  Net-new relative to the reference.
    """
    if mesh_cp.ndim != 1:
        raise ValueError("the context-parallel mesh must have exactly one dimension")
    group = mesh_cp.get_group()
    switched = 0
    for sub in module.modules():
        if isinstance(sub, FlashSdpa):
            sub.enable_context_parallel(group, mode=mode, layout=layout)
            switched += 1
    if switched == 0:
        raise ValueError(f"{type(module).__name__} contains no FlashSdpa attention kernel to parallelise")
    return switched
