from __future__ import annotations

import math

import torch
import torch.distributed as dist

_MAX_NDIM = 8


def gather(
    tensor: torch.Tensor, group: dist.ProcessGroup, group_dst: int, async_op: bool = False
) -> list[torch.Tensor] | None | tuple[list[torch.Tensor] | None, dist.Work]:
    """Same-shape gather to ``group_dst``; output buffers are allocated on the destination only."""
    sink = [torch.empty_like(tensor) for _ in range(group.size())] if group.rank() == group_dst else None
    work = dist.gather(tensor, sink, group=group, group_dst=group_dst, async_op=async_op)
    return (sink, work) if async_op else sink


def all_gather(
    tensor: torch.Tensor, group: dist.ProcessGroup, async_op: bool = False
) -> list[torch.Tensor] | tuple[list[torch.Tensor], dist.Work]:
    """Same-shape all-gather with allocated outputs."""
    sink = [torch.empty_like(tensor) for _ in range(group.size())]
    work = dist.all_gather(sink, tensor, group=group, async_op=async_op)
    return (sink, work) if async_op else sink


def _exchange_shapes(tensor: torch.Tensor, group: dist.ProcessGroup) -> list[tuple[int, ...]]:
    if tensor.ndim > _MAX_NDIM:
        raise ValueError(f"variadic collectives support at most {_MAX_NDIM} dims, got {tensor.ndim}")
    desc = torch.zeros(_MAX_NDIM + 1, dtype=torch.long)
    desc[0] = tensor.ndim
    for i, s in enumerate(tensor.shape):
        desc[i + 1] = s
    desc = desc.to(tensor.device)
    everyone = torch.empty(group.size() * (_MAX_NDIM + 1), dtype=torch.long, device=tensor.device)
    dist.all_gather_into_tensor(everyone, desc, group=group)
    rows = everyone.view(group.size(), _MAX_NDIM + 1).cpu().tolist()
    return [tuple(int(v) for v in row[1 : 1 + int(row[0])]) for row in rows]


def all_gather_variadic_shape(
    tensor: torch.Tensor, group: dist.ProcessGroup, async_op: bool = False
) -> list[torch.Tensor] | tuple[list[torch.Tensor], dist.Work]:
    """All-gather of tensors whose shapes differ per rank (the shape exchange itself is synchronous)."""
    shapes = _exchange_shapes(tensor, group)
    # one equal-size collective on a buffer padded to the largest shard (uneven list all-gathers are not portable
    # across backends); the results are views into the gathered buffer
    sizes = [math.prod(shape) for shape in shapes]
    width = max(max(sizes), 1)
    padded = tensor.new_zeros(width)
    padded[: tensor.numel()] = tensor.reshape(-1)
    gathered = tensor.new_empty(group.size() * width)
    work = dist.all_gather_into_tensor(gathered, padded, group=group, async_op=async_op)
    sink = [gathered[r * width : r * width + n].view(shape) for r, (n, shape) in enumerate(zip(sizes, shapes, strict=True))]
    return (sink, work) if async_op else sink


def gather_variadic_shape(tensor: torch.Tensor, group: dist.ProcessGroup, group_dst: int) -> list[torch.Tensor] | None:
    """Gather differently-shaped tensors on ``group_dst`` using point-to-point transfers."""
    shapes = _exchange_shapes(tensor, group)
    me = group.rank()
    if me != group_dst:
        dist.isend(tensor.contiguous(), group=group, group_dst=group_dst).wait()
        return None
    out: list[torch.Tensor] = []
    pending = []
    for src in range(group.size()):
        if src == me:
            out.append(tensor)
            continue
        buf = torch.empty(shapes[src], dtype=tensor.dtype, device=tensor.device)
        pending.append(dist.irecv(buf, group=group, group_src=src))
        out.append(buf)
    for work in pending:
        work.wait()
    return out
