from __future__ import annotations

import torch
from torch import nn

from d9d_b200.module.base import ModuleLateInit


class ClassificationHead(nn.Module, ModuleLateInit):
    """Dropout + bias-free projection to ``num_labels`` fp32 logits over (optionally mask-pooled) hidden states.

    Parity: reference ``d9d/module/block/head/classification.py:7-55``.
    """

    def __init__(self, hidden_size: int, num_labels: int, dropout: float):
        super().__init__()
        self.dropout = nn.Dropout(dropout)
        self.score = nn.Linear(hidden_size, num_labels, bias=False)

    def forward(self, hidden_states: torch.Tensor, pooling_mask: torch.Tensor | None) -> torch.Tensor:
        if pooling_mask is not None:
            hidden_states = hidden_states[pooling_mask == 1]
        return self.score(self.dropout(hidden_states)).float()

    def reset_parameters(self) -> None:
        self.score.reset_parameters()
