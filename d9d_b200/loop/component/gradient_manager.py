from __future__ import annotations

from collections.abc import Iterator
from contextlib import contextmanager

import torch
from torch.distributed.tensor import DTensor

from d9d_b200.core.dist_context import DistributedContext
from d9d_b200.internals.grad_sync import GradientSynchronizer
from d9d_b200.kernel._native import EXTERNAL_GRAD_OWNER_ATTR
from d9d_b200.loop.config import GradientManagerConfig
from d9d_b200.metric.impl.aggregation import WeightedMeanMetric

from .batch_maths import BatchMaths
from .model_stage_factory import TrackedModules


class GradientManager:
    """Gradient lifecycle of one optimizer step.

    Contract (reference ``gradient_manager.py:80-132``): the loss callback back-propagates ``loss * weight``;
    gradients are SUMmed over microbatches (in-place accumulation into the flat arenas) and over data-parallel
    replicas (bucketed all-reduce overlapped with backward); finally every gradient is multiplied by
    ``1 / sum_of_all_weights`` (weights summed over microbatches AND ranks), which makes the result an exact
    weighted mean over the global batch.
    """

    def __init__(self, dist_context: DistributedContext, tracked_modules: TrackedModules, batch_maths: BatchMaths,
                 config: GradientManagerConfig):
        self._ctx = dist_context
        self._modules = tracked_modules
        self._config = config
        self._loss = WeightedMeanMetric()
        self._loss.to(dist_context.current_device)
        self._num_backward_calls = batch_maths.num_backward_calls
        self._sync = GradientSynchronizer([list(m.parameters()) for m in tracked_modules.modules],
                                          bucket_size_mb=config.bucket_size_mb,
                                          require_accumulations=batch_maths.num_backward_calls)
        self._installed = False
        self._external_owners: list = []
        self._scale_sinks: list = []  # optimizers whose update kernel multiplies gradients by a device scalar
        self._pending_scale: torch.Tensor | None = None

    def bind_optimizer(self, optimizer: object) -> None:
        """Let the manager hand its gradient scale to the optimizer(s) instead of rewriting the gradients."""
        leaves = list(getattr(optimizer, "optimizers", [optimizer]))
        if self._config.fold_scaling_into_optimizer and leaves and all(hasattr(o, "grad_scale") and not hasattr(o, "set_grad_scale") for o in leaves):
            self._scale_sinks = leaves
            self._pending_scale = torch.ones(1, dtype=torch.float32, device=self._ctx.current_device)
            for o in leaves:
                o.grad_scale = self._pending_scale

    @property
    def pending_scale(self) -> torch.Tensor | None:
        """Device scalar that still has to be applied to every gradient (None: gradients are already scaled)."""
        return self._pending_scale

    def _apply_grad_dtype(self) -> None:
        if self._config.grad_dtype is None:
            return
        dtype = getattr(torch, self._config.grad_dtype)
        for module in self._modules.modules:
            for p in module.parameters():
                if p.requires_grad:
                    p.grad_dtype = dtype

    @property
    def synchronizer(self) -> GradientSynchronizer:
        return self._sync

    @contextmanager
    def install(self) -> Iterator[None]:
        """Set gradient dtypes, allocate the arenas / hooks; torn down on exit."""
        self._apply_grad_dtype()
        self._sync.bind()
        owners = {}
        for module in self._modules.modules:
            for p in module.parameters():
                owner = getattr(p, EXTERNAL_GRAD_OWNER_ATTR, None)
                if owner is not None:
                    owners[id(owner)] = owner
        self._external_owners = list(owners.values())  # optimizers that reduce / scale / zero their gradients themselves
        for owner in self._external_owners:
            if hasattr(owner, "set_required_accumulations"):  # lets them overlap their reduction with backward
                owner.set_required_accumulations(self._num_backward_calls)
        self._installed = True
        try:
            yield
        finally:
            self._installed = False
            self._sync.unbind()

    def add_loss_with_weight(self, loss: torch.Tensor, loss_weight: torch.Tensor) -> None:
        self._loss.update(loss, loss_weight)

    def _local_grads(self) -> list[torch.Tensor]:
        arenas = [a.buffer for a in self._sync.arenas]
        if arenas:
            return arenas  # every gradient aliases an arena: scale the flat buffers (few launches)
        grads = []
        for module in self._modules.modules:
            for p in module.parameters():
                if p.grad is not None and getattr(p, EXTERNAL_GRAD_OWNER_ATTR, None) is None:
                    grads.append(p.grad.to_local() if isinstance(p.grad, DTensor) else p.grad)
        return grads

    def sync_and_scale(self) -> None:
        if not self._installed:
            raise ValueError("You should bind the manager first.")
        self._sync.wait()
        if self._ctx.mesh_params.is_distributed:
            self._loss.sync(self._ctx)
        grads = self._local_grads()
        inv = 1.0 / self._loss.accumulated_weight
        if self._pending_scale is not None:
            self._pending_scale.copy_(inv.reshape(-1)[:1])  # consumed by the optimizer kernel (and the clipper)
        elif grads:
            torch._foreach_mul_(grads, inv)  # noqa: SLF001  tensor scalar: no host sync
        for owner in self._external_owners:
            owner.set_grad_scale(inv)  # applied inside the NVLink update kernel after the cross-replica reduction

    def compute_global_loss(self) -> torch.Tensor:
        return self._loss.compute()

    def zero_grad(self) -> None:
        self._sync.zero_grad()
        self._loss.reset()
