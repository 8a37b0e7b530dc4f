from __future__ import annotations

import os
import random

import torch

from d9d_b200.core.dist_context import REGULAR_DOMAIN, DistributedContext


def set_seeds(dist_context: DistributedContext, seed: int, distinct_seed_mesh_dim: str = "pp") -> None:
    """Seed python / numpy / torch with ``seed + rank along distinct_seed_mesh_dim``.

    Ranks that differ only in other mesh dims share the seed (so replicas initialise identically and EP/TP peers
    draw the same numbers); the DTensor RNG tracker is seeded consistently over those dims.
    """
    distributed = dist_context.mesh_params.is_distributed
    if distributed:
        regular = dist_context.mesh_for(REGULAR_DOMAIN)
        seed = (seed + regular[distinct_seed_mesh_dim].get_local_rank()) % 2**64
    dist_context.logger.info(f"Set seed {seed}")
    torch.manual_seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed % 2**32)
    random.seed(seed)
    try:
        import numpy as np

        np.random.seed(seed % 2**32)
    except ImportError:
        pass
    if distributed:
        regular = dist_context.mesh_for(REGULAR_DOMAIN)
        shared_dims = tuple(n for n in (regular.mesh_dim_names or ()) if n != distinct_seed_mesh_dim)
        if shared_dims:
            shared = regular[shared_dims]
            if shared.get_coordinate() is not None:
                try:
                    torch.distributed.tensor._random.manual_seed(seed % 2**63, shared)  # noqa: SLF001
                except Exception as exc:  # CPU meshes have no DTensor RNG tracker
                    dist_context.logger.debug(f"DTensor RNG not seeded: {exc}")
