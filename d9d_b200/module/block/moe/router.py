from __future__ import annotations

import dataclasses
from typing import Literal

import torch
from pydantic import BaseModel, PositiveInt
from torch import nn

from d9d_b200.kernel.router import route_topk
from d9d_b200.module.base import ModuleLateInit
from d9d_b200.module.block.linear import Linear


@dataclasses.dataclass(kw_only=True, slots=True)
class RoutingResult:
    selected_expert_indices: torch.Tensor  # [T, k] int64
    selected_probabilities: torch.Tensor  # [T, k] fp32


class RouterParameters(BaseModel):
    """Routing variants beyond the plain softmax top-k (DeepSeek-V2 / V3)."""

    score_function: Literal["softmax", "sigmoid"] = "softmax"
    enable_expert_bias: bool = False  # selection bias buffer for auxiliary-loss-free load balancing
    num_expert_groups: PositiveInt = 1  # group-limited routing: experts are split into this many groups ...
    topk_expert_groups: PositiveInt = 1  # ... of which only the best ones stay eligible
    routed_scaling_factor: float = 1.0


class TopKRouter(nn.Module, ModuleLateInit):
    """Linear gate -> fp32 softmax over all experts -> top-k (optionally biased selection) -> optional renorm
    (softmax / selection / renormalisation are one fused kernel on CUDA, ``ops/csrc/router.cu``).

    Parity: reference ``d9d/module/block/moe/router.py:24-107``.
    """

    def __init__(self, dim: int, num_experts: int, top_k: int, renormalize_probabilities: bool,
                 enable_expert_bias: bool = False, options: RouterParameters | None = None) -> None:
        super().__init__()
        options = options if options is not None else RouterParameters(enable_expert_bias=enable_expert_bias)
        self.gate = Linear(dim, num_experts, bias=False)
        self.expert_bias: nn.Buffer | None = (
            nn.Buffer(torch.empty(num_experts, dtype=torch.float32), persistent=True)
            if (enable_expert_bias or options.enable_expert_bias) else None
        )
        self._num_experts = num_experts
        self._top_k = top_k
        self._renormalize = renormalize_probabilities
        self._options = options

    def forward(self, hidden_states: torch.Tensor) -> RoutingResult:
        o = self._options
        chosen, chosen_p = route_topk(self.gate(hidden_states), self._top_k, self._renormalize, self.expert_bias,
                                      score_function=o.score_function, num_groups=o.num_expert_groups,
                                      topk_groups=o.topk_expert_groups, scaling_factor=o.routed_scaling_factor)
        return RoutingResult(selected_expert_indices=chosen, selected_probabilities=chosen_p)

    def reset_parameters(self) -> None:
        if self.expert_bias is not None:
            nn.init.zeros_(self.expert_bias)
        self.gate.reset_parameters()
