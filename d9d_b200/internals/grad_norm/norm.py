from __future__ import annotations

import math

import torch
import torch.distributed as dist
from torch import nn
from torch.autograd.profiler import record_function
from torch.distributed import DeviceMesh
from torch.distributed.tensor import DTensor

from .group import ParametersForNorm


def _local_grad(p: nn.Parameter) -> torch.Tensor:
    if p.grad is None:
        raise ValueError("None grad detected")
    return p.grad.to_local() if isinstance(p.grad, DTensor) else p.grad


def _reduce_op(norm_type: float) -> dist.ReduceOp.RedOpType:
    return dist.ReduceOp.MAX if math.isinf(norm_type) else dist.ReduceOp.SUM


def _device_of(groups: ParametersForNorm) -> torch.device:
    for key in groups:
        return key.device
    return torch.device("cuda" if torch.cuda.is_available() else "cpu")


def _local_norm_pow(params: list[nn.Parameter], norm_type: float, device: torch.device) -> torch.Tensor:
    if not params:
        return torch.zeros((), device=device)
    total = torch.nn.utils.get_total_norm([_local_grad(p) for p in params], norm_type=norm_type, foreach=True,
                                          error_if_nonfinite=False)
    return total if math.isinf(norm_type) else total**norm_type


def clip_grad_norm_distributed_(parameter_groups: ParametersForNorm, max_norm: float | None, norm_type: float,
                                pp_mesh: DeviceMesh | None, pending_scale: torch.Tensor | None = None) -> torch.Tensor:
    """Global gradient norm over every parallel dimension, then in-place clipping (tensor coefficient, no host sync).

    Sharded groups all-reduce their ``||g||^p`` over the mesh dim(s) they are sharded on (async, launched first);
    replicated groups contribute locally; finally one scalar all-reduce over the pipeline dimension.
    ``max_norm=None`` only measures.  ``pending_scale`` (device scalar) is a factor that has not been applied to the
    gradients yet: the norm is reported for the scaled gradients and the clip coefficient is multiplied into it instead
    of rewriting every gradient.  Parity: reference ``d9d/internals/grad_norm/norm.py:99-139``.
    """
    with record_function("Gradient Clipping"):
        device = _device_of(parameter_groups)
        partials: list[torch.Tensor] = []
        works: list[dist.Work] = []
        for key, params in parameter_groups.items():
            value = _local_norm_pow(params, norm_type, device)
            if key.shard_meshes is not None:
                # shards over several mesh dims (e.g. FSDP x tensor parallel) partition the tensor over the product of the
                # dims: reduce over all but the last dim now, the last one asynchronously like single-dim groups
                for mesh in key.shard_meshes[:-1]:
                    dist.all_reduce(value, op=_reduce_op(norm_type), group=mesh.get_group())
                works.append(dist.all_reduce(value, op=_reduce_op(norm_type), group=key.shard_meshes[-1].get_group(), async_op=True))
            partials.append(value)
        for w in works:
            w.wait()
        if partials:
            stacked = torch.stack([p.to(device) for p in partials])
            total_pow = stacked.max() if math.isinf(norm_type) else stacked.sum()
        else:
            total_pow = torch.zeros((), device=device)
        if pp_mesh is not None:
            dist.all_reduce(total_pow, op=_reduce_op(norm_type), group=pp_mesh.get_group())
        total = total_pow if math.isinf(norm_type) else total_pow ** (1.0 / norm_type)
        if pending_scale is not None:
            total = total * pending_scale.reshape(()).abs()
        if max_norm:
            coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
            if pending_scale is not None:
                pending_scale.mul_(coef)
            else:
                for params in parameter_groups.values():
                    torch._foreach_mul_([_local_grad(p) for p in params], coef)  # noqa: SLF001
        return total
