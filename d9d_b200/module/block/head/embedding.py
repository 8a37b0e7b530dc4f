from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from d9d_b200.module.base import ModuleLateInit


class EmbeddingHead(nn.Module, ModuleLateInit):
    """Optional projection + optional L2 normalisation (fp32) of mask-pooled hidden states.

    Parity: reference ``d9d/module/block/head/embedding.py:8-68``.
    """

    def __init__(self, hidden_size: int, embedding_dim: int | None, normalize: bool):
        super().__init__()
        self._normalize = normalize
        self.projection = nn.Linear(hidden_size, embedding_dim, bias=False) if embedding_dim is not None else None

    def forward(self, hidden_states: torch.Tensor, pooling_mask: torch.Tensor | None = None) -> torch.Tensor:
        if pooling_mask is not None:
            hidden_states = hidden_states[pooling_mask == 1]
        if self.projection is not None:
            hidden_states = self.projection(hidden_states)
        hidden_states = hidden_states.float()
        return F.normalize(hidden_states, p=2, dim=-1) if self._normalize else hidden_states

    def reset_parameters(self) -> None:
        if self.projection is not None:
            self.projection.reset_parameters()
