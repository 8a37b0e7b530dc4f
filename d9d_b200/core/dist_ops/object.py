from __future__ import annotations

from typing import TypeVar

import torch.distributed as dist

T = TypeVar("T")


def gather_object(obj: T, group: dist.ProcessGroup, group_dst: int) -> list[T] | None:
    """Gather picklable objects on ``group_dst`` (rank inside ``group``); other ranks get ``None``."""
    sink: list[T] | None = [None] * group.size() if group.rank() == group_dst else None  # type: ignore[list-item]
    dist.gather_object(obj, sink, group=group, group_dst=group_dst)
    return sink


def all_gather_object(obj: T, group: dist.ProcessGroup) -> list[T]:
    """Gather picklable objects on every rank."""
    sink: list[T] = [None] * group.size()  # type: ignore[list-item]
    dist.all_gather_object(sink, obj, group=group)
    return sink
