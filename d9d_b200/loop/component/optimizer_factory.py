from __future__ import annotations

from d9d_b200.core.dist_context import REGULAR_DOMAIN, DistributedContext
from d9d_b200.core.protocol import LRSchedulerProtocol, OptimizerProtocol
from d9d_b200.loop.control import (
    InitializeLRSchedulerContext,
    InitializeOptimizerStageContext,
    LRSchedulerProvider,
    OptimizerProvider,
)
from d9d_b200.pipelining.training import PipelinedLRScheduler, PipelinedOptimizer

from .model_stage_factory import TrackedModules
from .stepper import Stepper


class _FrozenStageOptimizer(OptimizerProtocol):
    """Stands in for the optimizer of a model stage without trainable parameters (e.g. the first pipeline stage when only a
    head on the last stage is tuned): torch optimizers refuse empty parameter lists."""

    param_groups: list = []

    def step(self) -> None: ...

    def zero_grad(self) -> None: ...

    def state_dict(self) -> dict:
        return {}

    def load_state_dict(self, state_dict: dict) -> None: ...


class _FrozenStageScheduler(LRSchedulerProtocol):
    def step(self) -> None: ...

    def state_dict(self) -> dict:
        return {}

    def load_state_dict(self, state_dict: dict) -> None: ...


class OptimizerFactory:
    """One optimizer + LR scheduler per local model stage, wrapped into the pipeline-aware aggregates."""

    def __init__(self, dist_context: DistributedContext, tracked_modules: TrackedModules, optimizer_provider: OptimizerProvider,
                 lr_scheduler_provider: LRSchedulerProvider, stepper: Stepper):
        self._ctx, self._modules = dist_context, tracked_modules
        self._opt_provider, self._lr_provider, self._stepper = optimizer_provider, lr_scheduler_provider, stepper

    def build_optimizer_and_scheduler(self) -> tuple[OptimizerProtocol, LRSchedulerProtocol]:
        optimizers, schedulers = [], []
        for module in self._modules.modules:
            if not any(p.requires_grad for p in module.parameters()):
                optimizers.append(_FrozenStageOptimizer())
                schedulers.append(_FrozenStageScheduler())
                continue
            opt = self._opt_provider(InitializeOptimizerStageContext(dist_context=self._ctx, model=module))
            optimizers.append(opt)
            schedulers.append(self._lr_provider(InitializeLRSchedulerContext(dist_context=self._ctx, total_steps=self._stepper.total_steps, optimizer=opt)))
        mesh_pp = self._ctx.mesh_for(REGULAR_DOMAIN)["pp"] if self._ctx.mesh_params.is_distributed else None
        return PipelinedOptimizer(mesh_pp=mesh_pp, optimizers=optimizers), PipelinedLRScheduler(mesh_pp=mesh_pp, schedulers=schedulers)
