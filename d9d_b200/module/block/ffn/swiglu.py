from __future__ import annotations

import torch
from torch import nn

from d9d_b200.kernel.swiglu import silu_mul
from d9d_b200.module.base import ModuleLateInit
from d9d_b200.module.block.linear import Linear


class SwiGLU(nn.Module, ModuleLateInit):
    """``down(silu(gate(x)) * up(x))`` (reference ``d9d/module/block/ffn/swiglu.py:8-49``)."""

    def __init__(self, hidden_size: int, intermediate_size: int, bias: bool = False):
        super().__init__()
        self.gate_proj = Linear(hidden_size, intermediate_size, bias=bias)
        self.up_proj = Linear(hidden_size, intermediate_size, bias=bias)
        self.down_proj = Linear(intermediate_size, hidden_size, bias=bias)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.down_proj(silu_mul(self.gate_proj(x), self.up_proj(x)))

    def reset_parameters(self) -> None:
        for proj in (self.gate_proj, self.up_proj, self.down_proj):
            proj.reset_parameters()
