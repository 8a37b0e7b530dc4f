from __future__ import annotations

from typing import Any

from torch import nn
from torch.distributed import DeviceMesh
from torch.distributed.tensor import DTensor, Placement, distribute_tensor
from torch.distributed.tensor.parallel import ParallelStyle

from d9d_b200.kernel._native import MAIN_PARAM_ATTR


class _LocalParameterView:
    """Forward pre/post hook pair that lends *local* tensors to the modules for the duration of a forward.

    Outside forward every parameter is a ``DTensor`` (so ``state_dict()`` / DCP / gradient-sync / grad-norm
    grouping see placements).  On entry each ``module._parameters[name]`` is replaced by
    ``param.to_local(grad_placements=...)`` — a view whose gradient flows back to the DTensor parameter — and the
    DTensor parameters are restored on exit (also when forward raises).  Re-entrant calls nest safely.
    """

    def __init__(self, slots: list[tuple[nn.Module, str]], grad_placement: tuple[Placement, ...]):
        self._slots = slots
        self._grad_placement = grad_placement
        self._depth = 0
        self._stash: list[Any] = []

    def enter(self, module: nn.Module, args: Any, kwargs: Any = None) -> None:
        self._depth += 1
        if self._depth > 1:
            return
        stash = []
        for owner, name in self._slots:
            param = owner._parameters[name]  # noqa: SLF001
            stash.append(param)
            if isinstance(param, DTensor):
                local = param.to_local(grad_placements=self._grad_placement)
                setattr(local, MAIN_PARAM_ATTR, param)  # lets wgrad kernels accumulate straight into param.grad
                owner._parameters[name] = local  # noqa: SLF001
        self._stash = stash

    def exit(self, module: nn.Module, args: Any, output: Any) -> None:
        self._depth -= 1
        if self._depth > 0:
            return
        for (owner, name), param in zip(self._slots, self._stash, strict=True):
            owner._parameters[name] = param  # noqa: SLF001
        self._stash = []


class ToLocalParallel(ParallelStyle):
    """Distribute parameters as DTensors with ``param_placement`` but run the module's math on local tensors;
    gradients arrive as DTensors with ``grad_placement`` (e.g. Replicate = "pending reduce").

    Parity: reference ``d9d/module/parallelism/style/to_local.py:9-77`` (which swaps ``__class__`` of every submodule
    on each forward; here only the parameter dict entries are swapped).
    """

    def __init__(self, param_placement: tuple[Placement, ...], grad_placement: tuple[Placement, ...],
                 skip_distributed: bool = False):
        self._param_placement = tuple(param_placement)
        self._grad_placement = tuple(grad_placement)
        self._skip_distributed = skip_distributed  # leave parameters another style (e.g. tensor parallel) already owns

    def _apply(self, module: nn.Module, device_mesh: DeviceMesh) -> nn.Module:
        slots: list[tuple[nn.Module, str]] = []
        for sub in module.modules():
            for name, param in list(sub._parameters.items()):  # noqa: SLF001
                if param is None:
                    continue
                if isinstance(param.data, DTensor):
                    if self._skip_distributed:
                        continue
                    raise ValueError(f"parameter {name} of {type(sub).__name__} is already distributed")
                dist_param = nn.Parameter(
                    distribute_tensor(param.data, device_mesh, self._param_placement, src_data_rank=None),
                    requires_grad=param.requires_grad,
                )
                sub.register_parameter(name, dist_param)
                slots.append((sub, name))
        view = _LocalParameterView(slots, self._grad_placement)
        module.register_forward_pre_hook(view.enter, with_kwargs=True)
        module.register_forward_hook(view.exit, always_call=True)
        return module
