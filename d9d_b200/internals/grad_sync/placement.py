from __future__ import annotations

from torch.distributed.tensor import Placement, Shard

try:  # FSDP2 over tensor-parallel parameters produces strided shards (a private placement type)
    from torch.distributed.tensor.placement_types import _StridedShard
except ImportError:  # pragma: no cover - older / newer torch layouts
    _StridedShard = ()  # type: ignore[assignment,misc]


def is_shard_placement(placement: Placement) -> bool:
    """True for ``Shard`` and for the strided shards of 2-D (FSDP x TP) parameters: nothing to reduce on that mesh dim."""
    return isinstance(placement, Shard) or (bool(_StridedShard) and isinstance(placement, _StridedShard))
