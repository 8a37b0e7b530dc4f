from __future__ import annotations

import os
from collections.abc import Iterable
from typing import Any

import torch
import torch.distributed as dist
from torch import nn
from torch.distributed.tensor import DTensor

from d9d_b200.internals.nvlink import SymmetricArena
from d9d_b200.kernel._native import EXTERNAL_GRAD_OWNER_ATTR, FUSED_WGRAD_ATTR, native_ops

_ALIGN = 8  # elements: 16-byte vectors of bf16 parameters / 32-byte pairs of fp32 gradient vectors


def _local(p: torch.Tensor) -> torch.Tensor:
    """The rank-local storage of a parameter (DTensor parameters keep it in ``_local_tensor``)."""
    return p._local_tensor if isinstance(p, DTensor) else p.data  # noqa: SLF001


class NvlinkShardedAdamW(torch.optim.Optimizer):
    """Data-parallel AdamW (stochastic rounding, bf16 parameters) whose gradient reduction, update and parameter
    broadcast are two kernels over NVLink / NVSwitch peer memory instead of an NCCL all-reduce plus a replicated
    optimizer step (kernels: ``ops/csrc/nvlink_optim.cu``).

    * all parameters live in one *symmetric* bf16 arena, all gradients in a symmetric fp32 arena
      (``param.data`` / ``param.grad`` become views; wgrad GEMMs accumulate into the arena in their epilogue);
    * rank ``r`` owns a contiguous ``1/world`` shard of both arenas and the AdamW moments for that shard only
      (optimizer state memory and optimizer work are divided by the number of replicas);
    * ``step()``: barrier → ``multimem.ld_reduce`` of the shard's gradients (summed inside the NVSwitch) with the
      sum of squares for clipping → scalar all-reduce of the norm → barrier → AdamW on the shard with the new bf16
      parameters written to every replica by ``multimem.st`` → zero the gradient arena → barrier.

    Mathematically this equals the reference's "all-reduce gradients, then every replica runs the same optimizer"
    (``d9d/internals/grad_sync`` + ``d9d/optim/stochastic/adamw.py``); gradients are SUMmed, ``grad_scale`` (a device
    scalar, e.g. ``1 / sum(loss weights)``) and the clip coefficient are applied inside the update kernel.
    """

    def __init__(self, params: Iterable[nn.Parameter], group: dist.ProcessGroup, lr: float, betas: tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 1e-2, state_dtype: torch.dtype = torch.bfloat16,
                 max_norm: float | None = None, seed: int = 0):
        params = [p for p in params if p.requires_grad]
        if not params:
            raise ValueError("NvlinkShardedAdamW needs at least one trainable parameter")
        if any(_local(p).dtype != torch.bfloat16 for p in params):
            raise ValueError("NvlinkShardedAdamW supports bf16 parameters only")
        super().__init__(params, {"lr": lr, "betas": betas, "eps": eps, "weight_decay": weight_decay})
        if len(self.param_groups) != 1:
            raise ValueError("NvlinkShardedAdamW supports a single parameter group")
        self._group = group
        self._world = group.size()
        self._rank = group.rank()
        self._max_norm = max_norm
        self._seed = seed
        self._step_count = 0
        device = _local(params[0]).device

        offsets, total = [], 0
        for p in params:
            offsets.append(total)
            total += (_local(p).numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        quantum = self._world * 1024
        total = (total + quantum - 1) // quantum * quantum
        self._numel = total
        self._shard = total // self._world
        self._begin = self._rank * self._shard

        self.param_arena = SymmetricArena(total, torch.bfloat16, device, group)
        self.grad_arena = SymmetricArena(total, torch.float32, device, group)
        self.param_arena.buffer.zero_()
        self.grad_arena.buffer.zero_()
        with torch.no_grad():
            for p, off in zip(params, offsets, strict=True):
                local = _local(p)
                view = self.param_arena.buffer[off : off + local.numel()].view(local.shape)
                view.copy_(local)
                gview = self.grad_arena.buffer[off : off + local.numel()].view(local.shape)
                p.grad_dtype = torch.float32
                if isinstance(p, DTensor):  # the parameter now lives in the symmetric arena; the DTensor wrapper stays valid
                    p._local_tensor.set_(view)  # noqa: SLF001
                    p.grad = DTensor.from_local(gview, p.device_mesh, p.placements, run_check=False)
                else:
                    p.data = view
                    p.grad = gview
                setattr(p, FUSED_WGRAD_ATTR, True)
                setattr(p, EXTERNAL_GRAD_OWNER_ATTR, self)
        self.exp_avg = torch.zeros(self._shard, dtype=state_dtype, device=device)
        self.exp_avg_sq = torch.zeros(self._shard, dtype=state_dtype, device=device)
        self._sumsq = torch.zeros(1, dtype=torch.float32, device=device)
        self._scale = torch.ones(1, dtype=torch.float32, device=device)
        self.grad_scale: torch.Tensor | None = None  # optional device scalar multiplied into every gradient
        self.last_grad_norm: torch.Tensor | None = None
        torch.cuda.synchronize(device)
        self.param_arena.barrier()

    @property
    def uses_multicast(self) -> bool:
        """NVSwitch multicast (``multimem``) path or explicit peer loads/stores.

        In-switch reduction divides the NVLink traffic of the reduce by ``world - 1`` and of the broadcast by the
        same factor, which pays from 4 replicas on; with 2 replicas plain peer loads/stores are faster (measured on
        2xB200: 8.6 ms vs 16.2 ms for the reduce of a 2.73 B-parameter arena).  ``D9D_NVLINK_MULTIMEM=0/1`` overrides.
        """
        available = self.param_arena.multicast_ptr != 0 and self.grad_arena.multicast_ptr != 0
        override = os.environ.get("D9D_NVLINK_MULTIMEM")
        if override is not None:
            return available and override == "1"
        return available and self._world >= 4

    state_is_materialized = True  # nothing is created lazily: checkpoint loading needs no warm-up step

    @property
    def max_norm(self) -> float | None:
        return self._max_norm

    @max_norm.setter
    def max_norm(self, value: float | None) -> None:
        self._max_norm = value

    def set_grad_scale(self, scale: torch.Tensor) -> None:
        """Device scalar multiplied into every (reduced) gradient inside the update kernel, e.g. ``1/sum(weights)``."""
        if self.grad_scale is None:
            self.grad_scale = torch.ones(1, dtype=torch.float32, device=self._sumsq.device)
        self.grad_scale.copy_(scale.reshape(-1)[:1])

    @torch.no_grad()
    def step(self, closure: Any = None) -> None:  # type: ignore[override]
        if closure is not None:
            raise ValueError("closures are not supported")
        ops = native_ops()
        group = self.param_groups[0]
        beta1, beta2 = group["betas"]
        self._step_count += 1
        begin, end = self._begin, self._begin + self._shard

        self.grad_arena.barrier()  # every replica finished its backward; its gradients are visible
        self._sumsq.zero_()
        multicast = self.uses_multicast
        ops.nvl_reduce_shard_(self.grad_arena.buffer, self.grad_arena.peer_ptrs_dev,
                              self.grad_arena.multicast_ptr if multicast else 0, begin, end, self._world, self._rank, self._sumsq)
        scale = self.grad_scale if self.grad_scale is not None else None
        if self._max_norm is not None:
            dist.all_reduce(self._sumsq, group=self._group)
            norm = self._sumsq.sqrt()
            if scale is not None:
                norm = norm * scale
            self.last_grad_norm = norm
            clip = torch.clamp(self._max_norm / (norm + 1e-6), max=1.0)
            self._scale.copy_(clip if scale is None else clip * scale)
            scale = self._scale
        self.grad_arena.barrier()  # all replicas have read my gradients: they may be overwritten / zeroed
        lr = group["lr"]
        ops.nvl_adamw_shard_(self.param_arena.buffer, self.grad_arena.buffer, self.exp_avg, self.exp_avg_sq,
                             self.param_arena.peer_ptrs_dev, self.param_arena.multicast_ptr if multicast else 0, begin, end,
                             self._world, self._rank,
                             float(lr), beta1, beta2, group["eps"], group["weight_decay"],
                             1.0 - beta1**self._step_count, 1.0 - beta2**self._step_count,
                             self._seed + 7919 * self._step_count, scale)
        self.grad_arena.buffer.zero_()
        self.param_arena.barrier()  # every shard owner's parameter writes have landed here before the next forward

    def zero_grad(self, set_to_none: bool = False) -> None:  # gradients are zeroed inside step()
        if set_to_none:
            raise ValueError("gradients alias the symmetric arena and cannot be set to None")
        self.grad_arena.buffer.zero_()

    def state_dict(self) -> dict[str, Any]:
        return {f"shard_{self._rank}_of_{self._world}": {"exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq},
                "step": self._step_count, "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:
        shard = state_dict[f"shard_{self._rank}_of_{self._world}"]
        self.exp_avg.copy_(shard["exp_avg"])
        self.exp_avg_sq.copy_(shard["exp_avg_sq"])
        self._step_count = int(state_dict["step"])
        for g, saved in zip(self.param_groups, state_dict["param_groups"], strict=True):
            g.update(saved)
