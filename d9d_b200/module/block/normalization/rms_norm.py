from __future__ import annotations

import torch
from torch import nn

from d9d_b200.kernel.normalization import rms_norm
from d9d_b200.module.base import ModuleLateInit


class RMSNorm(nn.Module, ModuleLateInit):
    """RMS normalisation over the last dim with a learnable scale (optionally zero-centred: ``weight + 1``).

    Parity: reference ``d9d/module/block/normalization/rms_norm.py:8-52``.
    """

    def __init__(self, hidden_size: int, eps: float = 1e-6, zero_centered: bool = False) -> None:
        super().__init__()
        self._eps = eps
        self._zero_centered = zero_centered
        self.weight = nn.Parameter(torch.empty(hidden_size))

    @property
    def eps(self) -> float:
        return self._eps

    @property
    def zero_centered(self) -> bool:
        return self._zero_centered

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return rms_norm(x, self.weight, eps=self._eps, zero_centered=self._zero_centered)

    def reset_parameters(self) -> None:
        with torch.no_grad():
            self.weight.fill_(0.0 if self._zero_centered else 1.0)

    def extra_repr(self) -> str:
        return f"{self.weight.shape[0]}, eps={self._eps}, zero_centered={self._zero_centered}"
