from __future__ import annotations

import dataclasses

import torch
from torch import nn

from d9d_b200.kernel.router import route_topk
from d9d_b200.module.base import ModuleLateInit
from d9d_b200.module.block.linear import Linear


@dataclasses.dataclass(kw_only=True, slots=True)
class RoutingResult:
    selected_expert_indices: torch.Tensor  # [T, k] int64
    selected_probabilities: torch.Tensor  # [T, k] fp32


class TopKRouter(nn.Module, ModuleLateInit):
    """Linear gate -> fp32 softmax over all experts -> top-k (optionally biased selection) -> optional renorm
    (softmax / selection / renormalisation are one fused kernel on CUDA, ``ops/csrc/router.cu``).

    Parity: reference ``d9d/module/block/moe/router.py:24-107``.
    """

    def __init__(self, dim: int, num_experts: int, top_k: int, renormalize_probabilities: bool,
                 enable_expert_bias: bool = False) -> None:
        super().__init__()
        self.gate = Linear(dim, num_experts, bias=False)
        self.expert_bias: nn.Buffer | None = (
            nn.Buffer(torch.empty(num_experts, dtype=torch.float32), persistent=True) if enable_expert_bias else None
        )
        self._num_experts = num_experts
        self._top_k = top_k
        self._renormalize = renormalize_probabilities

    def forward(self, hidden_states: torch.Tensor) -> RoutingResult:
        chosen, chosen_p = route_topk(self.gate(hidden_states), self._top_k, self._renormalize, self.expert_bias)
        return RoutingResult(selected_expert_indices=chosen, selected_probabilities=chosen_p)

    def reset_parameters(self) -> None:
        if self.expert_bias is not None:
            nn.init.zeros_(self.expert_bias)
        self.gate.reset_parameters()
