from __future__ import annotations

from collections.abc import Sequence

import torch
from torch import nn

from d9d_b200.module.base import ModuleLateInit


def vocab_ranges(split_vocab_size: dict[str, int], split_order: Sequence[str]) -> dict[str, tuple[int, int]]:
    """Global id range ``[start, end)`` of every named vocabulary split, in concatenation order."""
    ranges, cursor = {}, 0
    for name in split_order:
        size = split_vocab_size[name]
        ranges[name] = (cursor, cursor + size)
        cursor += size
    return ranges


class SplitTokenEmbeddings(nn.Module, ModuleLateInit):
    """Embedding table stored as several named tables (e.g. ``regular`` / ``special``) covering contiguous id ranges.

    Parity: reference ``d9d/module/block/embedding/shard_token_embedding.py:26-97`` (state-dict keys
    ``token_embedding.{split}.weight``).  Lookup = masked sum over the splits.
    """

    def __init__(self, split_vocab_size: dict[str, int], split_order: Sequence[str], hidden_size: int):
        super().__init__()
        self.token_embedding = nn.ModuleDict({name: nn.Embedding(size, hidden_size) for name, size in split_vocab_size.items()})
        self._ranges = vocab_ranges(split_vocab_size, split_order)
        self._split_order = tuple(split_order)
        self._hidden_size = hidden_size

    def forward(self, input_ids: torch.Tensor) -> torch.Tensor:
        if not self._split_order:
            raise ValueError("Embeddings are empty - perhaps no splits were configured")
        total: torch.Tensor | None = None
        for name in self._split_order:
            start, end = self._ranges[name]
            inside = (input_ids >= start) & (input_ids < end)
            local_ids = torch.where(inside, input_ids - start, 0)
            piece = self.token_embedding[name](local_ids) * inside.unsqueeze(-1)
            total = piece if total is None else total + piece
        assert total is not None
        return total

    def reset_parameters(self) -> None:
        for table in self.token_embedding.values():
            table.reset_parameters()
