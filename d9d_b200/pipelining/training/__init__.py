"""Optimizer / LR-scheduler wrappers aggregating the per-stage objects of one pipeline rank."""

from __future__ import annotations

from typing import Any

from torch.distributed import DeviceMesh

from d9d_b200.core.protocol import LRSchedulerProtocol, OptimizerProtocol


def _key(pp_rank: int, i: int) -> str:
    return f"pp_{pp_rank}_stage_{i}"


def ensure_optimizer_state_initialized(optimizer: Any) -> None:
    """Materialise lazily-created optimizer state (moments, step counters) without changing any parameter.

    ``torch.distributed.checkpoint`` loads *into the structure* returned by ``state_dict()``; an optimizer that has
    not stepped yet exposes an empty ``state`` and would silently resume with fresh moments (the reference has this
    problem).  A step with zero gradients, ``lr = 0`` and ``weight_decay = 0`` creates the state and is otherwise a
    no-op; the loaded values then overwrite it.
    """
    import torch

    groups = getattr(optimizer, "param_groups", None)
    if groups is None or getattr(optimizer, "state_is_materialized", False):
        return
    params = [p for g in groups for p in g["params"] if p.requires_grad]
    if not params or all(len(optimizer.state.get(p, {})) > 0 for p in params):
        return
    saved = []
    for g in groups:
        saved.append({k: g[k] for k in ("lr", "weight_decay") if k in g})
        if "lr" in g:
            g["lr"] = torch.zeros_like(g["lr"]) if isinstance(g["lr"], torch.Tensor) else 0.0
        if "weight_decay" in g:
            g["weight_decay"] = 0.0
    had_no_grad = []
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p, dtype=getattr(p, "grad_dtype", None) or p.dtype)
            had_no_grad.append(p)
    try:
        optimizer.step()
    finally:
        for p in had_no_grad:
            p.grad = None
        for g, old in zip(groups, saved, strict=True):
            g.update(old)


class PipelinedOptimizer(OptimizerProtocol):
    """State-dict keys ``pp_{pp_rank}_stage_{i}`` (reference ``pipelining/training/optimizer.py:23-30``)."""

    def __init__(self, mesh_pp: DeviceMesh | None, optimizers: list[OptimizerProtocol]):
        self._pp_rank = mesh_pp.get_local_rank() if mesh_pp is not None else 0
        self._optimizers = optimizers

    @property
    def optimizers(self) -> list[OptimizerProtocol]:
        return self._optimizers

    def state_dict(self) -> dict[str, Any]:
        for o in self._optimizers:
            ensure_optimizer_state_initialized(o)
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
