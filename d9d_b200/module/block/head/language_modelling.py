from __future__ import annotations

from collections.abc import Sequence

import torch
from torch import nn

from d9d_b200.kernel.cce import linear_cross_entropy
from d9d_b200.module.base import ModuleLateInit

LM_IGNORE_INDEX = -100
"""Label value skipped by the LM head (loss 0 at that position)."""


class SplitLanguageModellingHead(nn.Module, ModuleLateInit):
    """LM head stored as named vocabulary splits; returns the *per-token* NLL without materialising logits.

    Parity: reference ``d9d/module/block/head/language_modelling.py:10-76`` (keys ``lm_head.{split}.weight``).
    """

    def __init__(self, split_vocab_size: dict[str, int], split_order: Sequence[str], hidden_size: int):
        super().__init__()
        self.lm_head = nn.ModuleDict({name: nn.Linear(hidden_size, size, bias=False) for name, size in split_vocab_size.items()})
        self._split_order = tuple(split_order)
        self._hidden_size = hidden_size

    def classifier_weight(self) -> torch.Tensor:
        parts = [self.lm_head[name].weight for name in self._split_order]
        return parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)

    def classifier_blocks(self) -> list[torch.Tensor]:
        """The classifier as row blocks in vocabulary order (consumed in place by the fused linear-CE kernels)."""
        return [self.lm_head[name].weight for name in self._split_order]

    def forward(self, hidden_states: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        return linear_cross_entropy(hidden_states, self.classifier_blocks(), labels, ignore_index=LM_IGNORE_INDEX, reduction="none")

    def reset_parameters(self) -> None:
        for head in self.lm_head.values():
            head.reset_parameters()
