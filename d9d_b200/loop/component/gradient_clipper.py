from __future__ import annotations

from collections.abc import Iterator
from contextlib import contextmanager

from d9d_b200.core.dist_context import REGULAR_DOMAIN, DistributedContext
from d9d_b200.kernel._native import EXTERNAL_GRAD_OWNER_ATTR
from d9d_b200.internals.grad_norm import ParametersForNorm, clip_grad_norm_distributed_, group_parameters_for_norm
from d9d_b200.loop.config import GradientClippingConfig
from d9d_b200.tracker import BaseTrackerRun

from .model_stage_factory import TrackedModules
from .stepper import Stepper


class GradientClipper:
    """Global-norm clipping over every parallel dimension + periodic logging of ``l2_grad_norm_total``."""

    def __init__(self, dist_context: DistributedContext, tracked_modules: TrackedModules, config: GradientClippingConfig,
                 stepper: Stepper):
        self._ctx, self._modules, self._config, self._stepper = dist_context, tracked_modules, config, stepper
        self._groups: ParametersForNorm | None = None
        self._external_owners: list = []
        self._gradient_manager = None

    def bind_gradient_manager(self, gradient_manager: object) -> None:
        """Lets the clipper see (and extend) a gradient scale the manager has deferred to the optimizer."""
        self._gradient_manager = gradient_manager
        self.last_norm = None  # device tensor of the most recent global norm

    @contextmanager
    def install(self) -> Iterator[None]:
        params = [p for m in self._modules.modules for p in m.parameters()]
        owners = {id(o): o for o in (getattr(p, EXTERNAL_GRAD_OWNER_ATTR, None) for p in params) if o is not None}
        self._external_owners = list(owners.values())
        if self._external_owners:
            if any(p.requires_grad and getattr(p, EXTERNAL_GRAD_OWNER_ATTR, None) is None for p in params):
                raise NotImplementedError("mixing NVLink-reduced and bucket-reduced parameters is not supported yet")
            for owner in self._external_owners:  # the norm is only known after the in-kernel reduction: clip there
                owner.max_norm = self._config.max_norm
        self._groups = group_parameters_for_norm(p for p in params if getattr(p, EXTERNAL_GRAD_OWNER_ATTR, None) is None)
        try:
            yield
        finally:
            self._groups = None

    def clip_and_log(self, run: BaseTrackerRun) -> None:
        should_log = self._stepper.should_do_action(self._config.log_total_steps)
        if not self._config.max_norm and not should_log:
            return
        if self._groups is None:
            raise ValueError("Parameter groups are not configured")
        if self._external_owners:
            # clipping happens inside the optimizer's update kernel; log the norm of the previous step (device value)
            norm = self._external_owners[0].last_grad_norm
            self.last_norm = norm
            if should_log and norm is not None:
                run.scalar(name="l2_grad_norm_total", value=norm.item())
            return
        pp_mesh = self._ctx.mesh_for(REGULAR_DOMAIN)["pp"] if self._ctx.mesh_params.is_distributed else None
        pending = self._gradient_manager.pending_scale if self._gradient_manager is not None else None
        norm = clip_grad_norm_distributed_(parameter_groups=self._groups, max_norm=self._config.max_norm, norm_type=2.0, pp_mesh=pp_mesh,
                                           pending_scale=pending)
        self.last_norm = norm
        if should_log:
            run.scalar(name="l2_grad_norm_total", value=norm.item())
