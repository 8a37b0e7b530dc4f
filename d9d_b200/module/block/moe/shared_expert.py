from __future__ import annotations

import torch
from pydantic import BaseModel
from torch import nn

from d9d_b200.module.base import ModuleLateInit
from d9d_b200.module.block.ffn import SwiGLU
from d9d_b200.module.block.linear import Linear


class SharedExpertParameters(BaseModel):
    intermediate_size: int
    enable_gate: bool


class SharedSwiGLU(nn.Module, ModuleLateInit):
    """Always-on SwiGLU expert with an optional scalar sigmoid gate (reference ``moe/shared_expert.py:22-74``)."""

    def __init__(self, hidden_size: int, params: SharedExpertParameters):
        super().__init__()
        self.expert = SwiGLU(hidden_size=hidden_size, intermediate_size=params.intermediate_size)
        # N=1 is below the GEMM's 8-column granularity -> plain nn.Linear (a matrix-vector product)
        self.gate = nn.Linear(hidden_size, 1, bias=False) if params.enable_gate else None

    def forward(self, hidden_states: torch.Tensor) -> torch.Tensor:
        out = self.expert(hidden_states)
        if self.gate is not None:
            out = out * torch.sigmoid(self.gate(hidden_states))
        return out

    def reset_parameters(self) -> None:
        self.expert.reset_parameters()
        if self.gate is not None:
            self.gate.reset_parameters()
