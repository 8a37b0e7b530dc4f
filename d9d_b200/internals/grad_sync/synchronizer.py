from __future__ import annotations

import dataclasses

import torch
from torch import nn
from torch.distributed import DeviceMesh
from torch.distributed.tensor import DTensor

from d9d_b200.kernel._native import EXTERNAL_GRAD_OWNER_ATTR

from .bucket import AbstractGradientBucket, LocalGradientBucket, SyncGradientBucket, local_of
from .placement import is_shard_placement

_ARENA_ALIGN = 64  # elements; keeps every bucket / parameter slice 256-byte aligned for vectorised kernels


def find_reduce_mesh(data: DTensor) -> DeviceMesh | None:
    """Sub-mesh of the dims on which the tensor is ``Replicate`` (= where its gradient must be summed)."""
    dims = []
    for i, placement in enumerate(data.placements):
        if placement.is_replicate():
            dims.append(i)
        elif not is_shard_placement(placement):
            raise ValueError(f"Unknown grad placement: {placement}")
    if not dims:
        return None
    names = data.device_mesh.mesh_dim_names
    assert names is not None, "meshes must be named"
    return data.device_mesh[tuple(names[i] for i in dims)]


@dataclasses.dataclass(frozen=True)
class _ArenaKey:
    group_index: int
    reduce_mesh: DeviceMesh | None
    device: torch.device
    grad_dtype: torch.dtype


@dataclasses.dataclass
class GradientArena:
    """One flat gradient buffer shared by the buckets of a (param group, reduce mesh, device, dtype) class."""

    key: _ArenaKey
    buffer: torch.Tensor
    params: list[nn.Parameter]


class GradientSynchronizer:
    """Data-parallel gradient reduction with SUM semantics (averaging is the caller's job).

    * parameters are classed by ``(optimizer group, reduce sub-mesh, device, grad dtype)``;
    * every class gets ONE flat zero-initialised arena; buckets of at most ``bucket_size_mb`` are contiguous slices
      of it, filled in *reverse* parameter order (gradients become ready roughly back to front);
    * every ``param.grad`` aliases its slice, so accumulation over microbatches, the all-reduce, clipping and the
      optimizer all operate on the same flat memory.

    Parity: reference ``d9d/internals/grad_sync/synchronizer.py:175-251`` / ``bucket.py``.
    """

    def __init__(self, param_groups: list[list[nn.Parameter]], bucket_size_mb: int, require_accumulations: int):
        self._param_groups = param_groups
        self._bucket_bytes = int(bucket_size_mb) * 1024 * 1024
        self._require = require_accumulations
        self._stream: "torch.cuda.Stream | None" = None
        self._buckets: list[AbstractGradientBucket] = []
        self._arenas: list[GradientArena] = []

    # ------------------------------------------------------------------ planning
    def _classify(self) -> dict[_ArenaKey, list[nn.Parameter]]:
        classes: dict[_ArenaKey, list[nn.Parameter]] = {}
        for gi, group in enumerate(self._param_groups):
            for p in reversed(group):
                if not p.requires_grad or getattr(p, EXTERNAL_GRAD_OWNER_ATTR, None) is not None:
                    continue  # frozen, or reduced by an optimizer that owns its gradients (optim/nvlink)
                mesh = find_reduce_mesh(p.data) if isinstance(p.data, DTensor) else None
                grad_dtype = p.grad_dtype or p.dtype
                classes.setdefault(_ArenaKey(gi, mesh, p.device, grad_dtype), []).append(p)
        return classes

    def bind(self) -> None:
        """Allocate arenas, alias gradients, register hooks. Must precede the first backward."""
        classes = self._classify()
        on_cuda = any(k.device.type == "cuda" for k in classes)
        self._stream = torch.cuda.Stream() if on_cuda else None
        for key, params in classes.items():
            # split into buckets by byte budget (in parameter dtype bytes, like the reference)
            chunks: list[list[nn.Parameter]] = [[]]
            used = 0
            for p in params:
                nbytes = p.numel() * p.element_size()
                if used + nbytes >= self._bucket_bytes and chunks[-1]:
                    chunks.append([])
                    used = 0
                chunks[-1].append(p)
                used += nbytes
            sizes = [sum(local_of(p.data).numel() for p in chunk) for chunk in chunks]
            starts, cursor = [], 0
            for n in sizes:
                starts.append(cursor)
                cursor += (n + _ARENA_ALIGN - 1) // _ARENA_ALIGN * _ARENA_ALIGN
            arena = torch.zeros(max(cursor, 1), dtype=key.grad_dtype, device=key.device)
            self._arenas.append(GradientArena(key, arena, params))
            for chunk, start in zip(chunks, starts, strict=True):
                if key.reduce_mesh is None:
                    bucket: AbstractGradientBucket = LocalGradientBucket(chunk, arena, start)
                else:
                    bucket = SyncGradientBucket(chunk, arena, start, self._require, key.reduce_mesh,
                                                self._stream if key.device.type == "cuda" else None)
                bucket.bind()
                self._buckets.append(bucket)

    def unbind(self) -> None:
        for b in self._buckets:
            b.unbind()
        self._buckets = []
        self._arenas = []
        self._stream = None

    # ------------------------------------------------------------------ step-time API
    def wait(self) -> None:
        """Make the compute stream wait for all outstanding reductions; raises if a bucket never fired."""
        if self._stream is not None:
            torch.cuda.current_stream().wait_stream(self._stream)
        for b in self._buckets:
            b.mark_sync()

    def zero_grad(self) -> None:
        for b in self._buckets:
            b.zero_grad()

    @property
    def buckets(self) -> list[AbstractGradientBucket]:
        return self._buckets

    @property
    def arenas(self) -> list[GradientArena]:
        """Flat gradient buffers (for fused clipping / optimizer kernels)."""
        return self._arenas
