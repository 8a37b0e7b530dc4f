from __future__ import annotations

import enum
import functools

import torch
import torch.distributed as dist


class ContextParallelLayout(enum.StrEnum):
    """Which tokens of a sequence a context-parallel rank holds.

    ``contiguous``: rank ``r`` holds the ``r``-th of ``world`` equal chunks.  With causal attention the last rank then
    does ``world`` times the work of the first.
    ``zigzag``: the sequence is cut into ``2 * world`` chunks and rank ``r`` holds chunks ``r`` and ``2 * world - 1 - r``,
    which gives every rank the same number of visible (query, key) pairs under a causal mask.
    """

    contiguous = "contiguous"
    zigzag = "zigzag"


@functools.lru_cache(maxsize=256)
def _indices(seq_len: int, world: int, rank: int, layout: str) -> torch.Tensor:
    if layout == ContextParallelLayout.contiguous:
        if seq_len % world != 0:
            raise ValueError(f"sequence length {seq_len} is not divisible by the context-parallel degree {world}")
        chunk = seq_len // world
        return torch.arange(rank * chunk, (rank + 1) * chunk)
    if layout == ContextParallelLayout.zigzag:
        if seq_len % (2 * world) != 0:
            raise ValueError(f"sequence length {seq_len} is not divisible by 2 * context-parallel degree ({2 * world})")
        chunk = seq_len // (2 * world)
        head = torch.arange(rank * chunk, (rank + 1) * chunk)
        tail = torch.arange((2 * world - 1 - rank) * chunk, (2 * world - rank) * chunk)
        return torch.cat([head, tail])
    raise ValueError(f"unknown context-parallel layout {layout!r}")


def local_sequence_indices(seq_len: int, world: int, rank: int, layout: ContextParallelLayout = ContextParallelLayout.zigzag,
                           device: torch.device | str | None = None) -> torch.Tensor:
    """Global positions (ascending within each chunk) of the tokens held by ``rank``: int64 ``[seq_len / world]``."""
    idx = _indices(int(seq_len), int(world), int(rank), str(layout))
    return idx.to(device) if device is not None else idx


def shard_sequence(x: torch.Tensor, dim: int, world: int, rank: int,
                   layout: ContextParallelLayout = ContextParallelLayout.zigzag) -> torch.Tensor:
    """The slice of ``x`` along ``dim`` that ``rank`` holds."""
    return x.index_select(dim, local_sequence_indices(x.shape[dim], world, rank, layout, x.device))


def gather_sequence(x_local: torch.Tensor, dim: int, group: dist.ProcessGroup,
                    layout: ContextParallelLayout = ContextParallelLayout.zigzag) -> torch.Tensor:
    """Inverse of :func:`shard_sequence` (all-gather + un-permutation); not differentiable - for outputs and tests."""
    world = group.size()
    parts = [torch.empty_like(x_local) for _ in range(world)]
    dist.all_gather(parts, x_local.contiguous(), group=group)
    seq_len = x_local.shape[dim] * world
    full = torch.cat(parts, dim=dim)
    order = torch.cat([local_sequence_indices(seq_len, world, r, layout, x_local.device) for r in range(world)])
    inverse = torch.empty_like(order)
    inverse[order] = torch.arange(seq_len, device=order.device)
    return full.index_select(dim, inverse)
