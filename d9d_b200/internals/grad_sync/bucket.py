from __future__ import annotations

import abc

import torch
import torch.distributed as dist
from torch import nn
from torch.autograd.profiler import record_function
from torch.distributed import DeviceMesh
from torch.distributed.tensor import DTensor
from torch.utils.hooks import RemovableHandle

from d9d_b200.kernel._native import FUSED_WGRAD_ATTR


def local_of(t: torch.Tensor) -> torch.Tensor:
    return t.to_local() if isinstance(t, DTensor) else t


def grad_like(param_data: torch.Tensor, local_grad: torch.Tensor) -> torch.Tensor:
    """Wrap a local gradient view with the parameter's DTensor metadata (or return it as is for plain params)."""
    if isinstance(param_data, DTensor):
        return DTensor.from_local(local_grad, device_mesh=param_data.device_mesh, placements=param_data.placements,
                                  run_check=False, shape=param_data.shape, stride=param_data.stride())
    return local_grad


class AbstractGradientBucket(abc.ABC):
    """A set of parameters whose gradients share one flat buffer and one synchronisation lifecycle."""

    @abc.abstractmethod
    def bind(self) -> None: ...

    @abc.abstractmethod
    def unbind(self) -> None: ...

    @abc.abstractmethod
    def zero_grad(self) -> None: ...

    @abc.abstractmethod
    def mark_sync(self) -> None: ...


class _FlatBucket(AbstractGradientBucket):
    """Owns a slice ``[start, stop)`` of a flat gradient arena; every ``param.grad`` is a view into it."""

    def __init__(self, params: list[nn.Parameter], arena: torch.Tensor, start: int):
        self._params = params
        self._arena = arena
        self._start = start
        self._stop = start + sum(local_of(p.data).numel() for p in params)
        self._bound = False

    @property
    def buffer(self) -> torch.Tensor:
        return self._arena[self._start : self._stop]

    @property
    def parameters(self) -> list[nn.Parameter]:
        return self._params

    @torch.no_grad()
    def _alias_grads(self) -> None:
        off = self._start
        for p in self._params:
            local = local_of(p.data)
            view = self._arena[off : off + local.numel()].view(local.shape)
            p.grad = grad_like(p.data, view)
            off += local.numel()

    @torch.no_grad()
    def bind(self) -> None:
        self._alias_grads()
        for p in self._params:
            setattr(p, FUSED_WGRAD_ATTR, True)  # wgrad GEMMs may accumulate into the arena in their epilogue
        self._bound = True

    @torch.no_grad()
    def unbind(self) -> None:
        for p in self._params:
            p.grad = None
            setattr(p, FUSED_WGRAD_ATTR, False)
        self._bound = False

    @torch.no_grad()
    def zero_grad(self) -> None:
        if not self._bound:
            raise ValueError("Buffer is not initialized")
        self.buffer.zero_()


class LocalGradientBucket(_FlatBucket):
    """Parameters with no Replicate mesh dim: nothing to reduce (grads still live in the arena so that clipping and
    the optimizer can work on flat memory)."""

    def mark_sync(self) -> None:
        return None


class SyncGradientBucket(_FlatBucket):
    """Gradients that must be SUM-reduced over the mesh dims where the parameter is replicated.

    A per-parameter ``post_accumulate_grad_hook`` counts accumulations in O(1); when every parameter of the bucket
    has been accumulated ``require_accumulations`` times the bucket issues its all-reduce(s) on the communication
    stream (side stream on CUDA, inline on CPU) so they overlap the rest of backward.
    """

    def __init__(self, params: list[nn.Parameter], arena: torch.Tensor, start: int, require_accumulations: int,
                 reduce_mesh: DeviceMesh, communicate_stream: "torch.cuda.Stream | None"):
        if not all(isinstance(p.data, DTensor) for p in params):
            raise ValueError("All parameters passed in synchronizable bucket should contain DTensor data")
        super().__init__(params, arena, start)
        self._require = require_accumulations
        # innermost mesh dim first (fastest links / smallest groups first)
        self._groups: list[dist.ProcessGroup] = list(reduce_mesh.get_all_groups())[::-1]
        self._stream = communicate_stream
        self._counts: dict[int, int] = {}
        self._pending = 0
        self._hooks: list[RemovableHandle] = []
        self._ready = False
        self._reset_counters()

    def _reset_counters(self) -> None:
        self._counts = {id(p): 0 for p in self._params}
        self._pending = len(self._params)

    @torch.no_grad()
    def _on_grad_accumulated(self, param: nn.Parameter) -> None:
        key = id(param)
        self._counts[key] += 1
        if self._counts[key] == self._require:
            self._pending -= 1
        elif self._counts[key] > self._require:
            raise ValueError("Tried to accumulate, but synchronization was not performed")
        if self._pending != 0:
            return
        if self._ready:
            raise ValueError("Tried to accumulate, but synchronization was not performed")
        with record_function("Gradient Sync"):
            buf = self.buffer
            if self._stream is not None:
                self._stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(self._stream):
                    for group in self._groups:
                        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
            else:
                for group in self._groups:
                    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        self._ready = True

    @torch.no_grad()
    def bind(self) -> None:
        super().bind()
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad_accumulated) for p in self._params]

    @torch.no_grad()
    def unbind(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []
        super().unbind()

    @torch.no_grad()
    def zero_grad(self) -> None:
        super().zero_grad()
        self._reset_counters()

    def mark_sync(self) -> None:
        if not self._ready:
            raise ValueError("This bucket is not ready for sync.")
        self._ready = False
