"""Optimizer / LR-scheduler wrappers aggregating the per-stage objects of one pipeline rank."""

from __future__ import annotations

from typing import Any

from torch.distributed import DeviceMesh

from d9d_b200.core.protocol import LRSchedulerProtocol, OptimizerProtocol


def _key(pp_rank: int, i: int) -> str:
    return f"pp_{pp_rank}_stage_{i}"


class PipelinedOptimizer(OptimizerProtocol):
    """State-dict keys ``pp_{pp_rank}_stage_{i}`` (reference ``pipelining/training/optimizer.py:23-30``)."""

    def __init__(self, mesh_pp: DeviceMesh | None, optimizers: list[OptimizerProtocol]):
        self._pp_rank = mesh_pp.get_local_rank() if mesh_pp is not None else 0
        self._optimizers = optimizers

    @property
    def optimizers(self) -> list[OptimizerProtocol]:
        return self._optimizers

    def state_dict(self) -> dict[str, Any]:
        return {_key(self._pp_rank, i): o.state_dict() for i, o in enumerate(self._optimizers)}

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:
        for i, o in enumerate(self._optimizers):
            o.load_state_dict(state_dict[_key(self._pp_rank, i)])

    def step(self) -> None:
        for o in self._optimizers:
            o.step()

    def zero_grad(self) -> None:
        for o in self._optimizers:
            o.zero_grad()


class PipelinedLRScheduler(LRSchedulerProtocol):
    def __init__(self, mesh_pp: DeviceMesh | None, schedulers: list[LRSchedulerProtocol]):
        self._pp_rank = mesh_pp.get_local_rank() if mesh_pp is not None else 0
        self._schedulers = schedulers

    def state_dict(self) -> dict[str, Any]:
        return {_key(self._pp_rank, i): s.state_dict() for i, s in enumerate(self._schedulers)}

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:
        for i, s in enumerate(self._schedulers):
            s.load_state_dict(state_dict[_key(self._pp_rank, i)])

    def step(self) -> None:
        for s in self._schedulers:
            s.step()


__all__ = ["PipelinedLRScheduler", "PipelinedOptimizer"]
