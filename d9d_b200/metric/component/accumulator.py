from __future__ import annotations

import enum
from collections.abc import Iterable
from typing import Any

import torch
import torch.distributed as dist
from torch.distributed.checkpoint.stateful import Stateful


class MetricReduceOp(enum.StrEnum):
    """No ``avg`` on purpose: averaging partial accumulators is not a safe way to combine metric state."""

    sum = "sum"
    max = "max"
    min = "min"


_TORCH_OP = {MetricReduceOp.sum: dist.ReduceOp.SUM, MetricReduceOp.max: dist.ReduceOp.MAX, MetricReduceOp.min: dist.ReduceOp.MIN}


class MetricAccumulator(Stateful):
    """A tensor accumulated locally every step plus a synchronised copy filled by an all-reduce over the world.

    Parity: reference ``d9d/metric/component/accumulator.py:42-141``.
    """

    def __init__(self, initial_value: torch.Tensor, reduce_op: MetricReduceOp = MetricReduceOp.sum):
        self._initial = initial_value.clone()
        self._local = initial_value.clone()
        self._synchronized = initial_value.clone()
        self._reduce_op = reduce_op
        self._is_synchronized = False

    @property
    def reduce_op(self) -> MetricReduceOp:
        return self._reduce_op

    def update(self, value: torch.Tensor | float | bool) -> None:
        if self._reduce_op == MetricReduceOp.sum:
            self._local.add_(value)
        else:
            if not isinstance(value, torch.Tensor):
                raise ValueError(f"Non-tensor inputs are not supported for `{self._reduce_op.value}` reduce op")
            pick = torch.maximum if self._reduce_op == MetricReduceOp.max else torch.minimum
            self._local.copy_(pick(self._local, value))
        self._is_synchronized = False

    def sync(self) -> None:
        self._synchronized.copy_(self._local)
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(self._synchronized, op=_TORCH_OP[self._reduce_op])
        self._is_synchronized = True

    @property
    def value(self) -> torch.Tensor:
        return self._synchronized if self._is_synchronized else self._local

    def reset(self) -> None:
        self._local.copy_(self._initial)
        self._is_synchronized = False

    def to(self, device: str | torch.device | int) -> None:
        self._initial = self._initial.to(device)
        self._local = self._local.to(device)
        self._synchronized = self._synchronized.to(device)

    @staticmethod
    def _rank_key() -> str:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        return f"local.rank{rank}"

    def state_dict(self) -> dict[str, Any]:
        """Checkpoint view.  No collective runs in here (``state_dict`` may be called by a subset of the ranks, or a different
        number of times per rank): every rank stores ITS OWN partial value under a rank-qualified key - the same convention
        as the data-loader and task states - so a resumed job continues each rank's accumulation exactly."""
        return {self._rank_key(): self._local.clone(), "synchronized": self._synchronized.clone(), "is_synchronized": self._is_synchronized}

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:
        """Restores this rank's partial value.  Checkpoints written by the earlier format (one rank-independent ``local``
        key holding the global reduction) are still understood: the total goes to rank 0, the other ranks restart from the
        identity (max / min are idempotent, every rank takes the value)."""
        key = self._rank_key()
        if key in state_dict:
            local = state_dict[key].detach().clone().to(self._local.device)
        elif "local" in state_dict:
            local = state_dict["local"].detach().clone().to(self._local.device)
            distributed = dist.is_available() and dist.is_initialized()
            if distributed and self._reduce_op == MetricReduceOp.sum and dist.get_rank() != 0:
                local = self._initial.clone()
        else:
            raise KeyError(f"metric accumulator state has no entry for this rank ({key}); it was saved by a different world size")
        self._local = local
        self._synchronized = state_dict["synchronized"].detach().clone().to(self._synchronized.device)
        self._is_synchronized = bool(state_dict["is_synchronized"])


def sync_accumulators(accumulators: Iterable[MetricAccumulator]) -> None:
    """Synchronise many accumulators with ONE collective per (reduce op, dtype, device) by flattening them into a
    single buffer — metric tensors are tiny, so launch latency, not bandwidth, is what matters."""
    accs = list(accumulators)
    if not (dist.is_available() and dist.is_initialized()):
        for a in accs:
            a.sync()
        return
    groups: dict[tuple, list[MetricAccumulator]] = {}
    for a in accs:
        groups.setdefault((a.reduce_op, a._local.dtype, a._local.device), []).append(a)  # noqa: SLF001
    for (op, _dtype, _device), members in groups.items():
        flat = torch.cat([m._local.reshape(-1) for m in members])  # noqa: SLF001
        dist.all_reduce(flat, op=_TORCH_OP[op])
        off = 0
        for m in members:
            n = m._local.numel()  # noqa: SLF001
            m._synchronized.copy_(flat[off : off + n].view_as(m._synchronized))  # noqa: SLF001
            m._is_synchronized = True  # noqa: SLF001
            off += n
