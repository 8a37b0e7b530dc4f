from __future__ import annotations

import torch
from torch.utils.data import Dataset


class SyntheticTokenDataset(Dataset[dict[str, torch.Tensor]]):
    """Deterministic random-token causal-LM samples of fixed length (``input_ids``, shifted ``labels``,
    ``position_ids``); sample ``i`` only depends on ``(seed, i)``.  Used by benchmarks and smoke tests."""

    def __init__(self, num_samples: int, seq_len: int, vocab_size: int, seed: int = 0, learnable: bool = False):
        self._n, self._s, self._v, self._seed = num_samples, seq_len, vocab_size, seed
        self._learnable = learnable  # next token is a fixed function of the current one (so a model can fit it)

    def __len__(self) -> int:
        return self._n

    def sort_key(self, index: int) -> int:
        return self._s

    def __getitem__(self, index: int) -> dict[str, torch.Tensor]:
        g = torch.Generator().manual_seed(self._seed * 1_000_003 + index)
        if self._learnable:
            start = int(torch.randint(0, self._v, (1,), generator=g))
            tokens = (start + 7 * torch.arange(self._s + 1)) % self._v
        else:
            tokens = torch.randint(0, self._v, (self._s + 1,), generator=g)
        return {"input_ids": tokens[:-1], "labels": tokens[1:], "position_ids": torch.arange(self._s)}

    @staticmethod
    def collate(batch: list[dict[str, torch.Tensor]]) -> dict[str, torch.Tensor]:
        return {k: torch.stack([b[k] for b in batch]) for k in batch[0]}
