from __future__ import annotations

import abc
from collections.abc import Iterable
from typing import Annotated, Literal

import torch
from pydantic import BaseModel, Field
from torch import nn
from torch.optim import SGD, Adam, AdamW, Optimizer

from d9d_b200.loop.control import InitializeOptimizerStageContext, OptimizerProvider
from d9d_b200.optim.stochastic import StochasticAdamW


def _fused_ok(params: list[nn.Parameter]) -> bool:
    return bool(params) and all(p.is_cuda for p in params)


class BaseAutoOptimizerConfig(BaseModel, abc.ABC):
    @abc.abstractmethod
    def build(self, params: Iterable[nn.Parameter]) -> Optimizer: ...


class StochasticAdamWOptimizerConfig(BaseAutoOptimizerConfig):
    name: Literal["stochastic_adamw"] = "stochastic_adamw"
    lr: float
    betas: tuple[float, float] = (0.9, 0.999)
    eps: float = 1e-8
    weight_decay: float = 1e-2
    state_dtype: str

    def build(self, params: Iterable[nn.Parameter]) -> Optimizer:
        return StochasticAdamW(params=params, lr=self.lr, betas=self.betas, eps=self.eps, weight_decay=self.weight_decay,
                               state_dtype=getattr(torch, self.state_dtype))


class AdamWOptimizerConfig(BaseAutoOptimizerConfig):
    name: Literal["adamw"] = "adamw"
    lr: float
    betas: tuple[float, float] = (0.9, 0.999)
    eps: float = 1e-8
    weight_decay: float = 1e-2
    amsgrad: bool = False
    maximize: bool = False

    def build(self, params: Iterable[nn.Parameter]) -> Optimizer:
        params = list(params)
        return AdamW(params=params, lr=self.lr, betas=self.betas, eps=self.eps, weight_decay=self.weight_decay,
                     amsgrad=self.amsgrad, maximize=self.maximize, fused=_fused_ok(params))


class AdamOptimizerConfig(BaseAutoOptimizerConfig):
    name: Literal["adam"] = "adam"
    lr: float
    betas: tuple[float, float] = (0.9, 0.999)
    eps: float = 1e-8
    weight_decay: float = 1e-2
    decoupled_weight_decay: bool = False
    amsgrad: bool = False
    maximize: bool = False

    def build(self, params: Iterable[nn.Parameter]) -> Optimizer:
        params = list(params)
        return Adam(params=params, lr=self.lr, betas=self.betas, eps=self.eps, weight_decay=self.weight_decay,
                    decoupled_weight_decay=self.decoupled_weight_decay, amsgrad=self.amsgrad, maximize=self.maximize,
                    fused=_fused_ok(params))


class SGDOptimizerConfig(BaseAutoOptimizerConfig):
    name: Literal["sgd"] = "sgd"
    lr: float
    momentum: float = 0
    dampening: float = 0
    weight_decay: float = 0
    nesterov: bool = False
    maximize: bool = False

    def build(self, params: Iterable[nn.Parameter]) -> Optimizer:
        params = list(params)
        return SGD(params, lr=self.lr, momentum=self.momentum, dampening=self.dampening, weight_decay=self.weight_decay,
                   nesterov=self.nesterov, maximize=self.maximize, fused=_fused_ok(params))


class NvlinkShardedAdamWOptimizerConfig(BaseAutoOptimizerConfig):
    """Stochastic-rounding AdamW whose data-parallel gradient reduction, update and parameter broadcast are NVLink
    peer-memory kernels (``d9d_b200.optim.nvlink``).  Pure data-parallel-replicate jobs only."""

    name: Literal["nvlink_sharded_adamw"] = "nvlink_sharded_adamw"
    lr: float
    betas: tuple[float, float] = (0.9, 0.999)
    eps: float = 1e-8
    weight_decay: float = 1e-2
    state_dtype: str = "bfloat16"

    def build(self, params: Iterable[nn.Parameter]) -> Optimizer:
        raise ValueError("nvlink_sharded_adamw needs the distributed context: use AutoOptimizerProvider")

    def build_distributed(self, params: Iterable[nn.Parameter], context: InitializeOptimizerStageContext) -> Optimizer:
        import torch.distributed as dist

        from d9d_b200.optim.nvlink import NvlinkShardedAdamW

        mesh = context.dist_context.mesh_params
        if mesh.world_size != mesh.data_parallel_replicate or mesh.has_expert_parallel or not mesh.is_distributed:
            raise ValueError("nvlink_sharded_adamw supports pure data-parallel-replicate meshes (dp_replicate == world size > 1)")
        return NvlinkShardedAdamW(params, dist.group.WORLD, lr=self.lr, betas=self.betas, eps=self.eps,
                                  weight_decay=self.weight_decay, state_dtype=getattr(torch, self.state_dtype))


AutoOptimizerConfig = Annotated[
    StochasticAdamWOptimizerConfig | AdamWOptimizerConfig | AdamOptimizerConfig | SGDOptimizerConfig | NvlinkShardedAdamWOptimizerConfig,
    Field(discriminator="name"),
]


class AutoOptimizerProvider(OptimizerProvider):
    def __init__(self, config: BaseAutoOptimizerConfig):
        self._config = config

    def __call__(self, context: InitializeOptimizerStageContext) -> Optimizer:
        params = (p for p in context.model.parameters() if p.requires_grad)
        if isinstance(self._config, NvlinkShardedAdamWOptimizerConfig):
            return self._config.build_distributed(params, context)
        return self._config.build(params)
