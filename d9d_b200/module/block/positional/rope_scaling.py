"""RoPE frequency scalings (reference ``d9d/module/block/positional/rope_scaling.py:11-137``)."""

from __future__ import annotations

import abc
import math

import torch


def base_inverse_frequencies(rope_base: float, dim: int) -> torch.Tensor:
    exponents = torch.arange(0, dim, 2, dtype=torch.float32) / dim
    return torch.pow(torch.tensor(float(rope_base), dtype=torch.float32), -exponents)


class RopeScaling(abc.ABC):
    """Strategy producing the per-pair inverse frequencies and the attention magnitude scale."""

    @abc.abstractmethod
    def inverse_frequencies(self, rope_base: int, head_dim: int) -> torch.Tensor: ...

    @property
    def attention_mscale(self) -> float:
        return 1.0


class NoRopeScaling(RopeScaling):
    def inverse_frequencies(self, rope_base: int, head_dim: int) -> torch.Tensor:
        return base_inverse_frequencies(rope_base, head_dim)


class LinearRopeScaling(RopeScaling):
    """Position interpolation: every frequency divided by ``factor``."""

    def __init__(self, factor: float) -> None:
        self._factor = factor

    def inverse_frequencies(self, rope_base: int, head_dim: int) -> torch.Tensor:
        return base_inverse_frequencies(rope_base, head_dim) / self._factor


class NtkRopeScaling(RopeScaling):
    """NTK-aware scaling: enlarge the base so low frequencies stretch and high ones stay."""

    def __init__(self, factor: float) -> None:
        self._factor = factor

    def inverse_frequencies(self, rope_base: int, head_dim: int) -> torch.Tensor:
        scaled_base = float(rope_base) * self._factor ** (head_dim / (head_dim - 2))
        return base_inverse_frequencies(scaled_base, head_dim)


class YarnRopeScaling(RopeScaling):
    """YaRN: blend interpolated and original frequencies with a linear ramp between two rotation counts."""

    def __init__(self, factor: float, beta_fast: float, beta_slow: float, original_max_position_embeddings: int) -> None:
        if beta_fast <= beta_slow:
            raise ValueError(f"beta_fast ({beta_fast}) must exceed beta_slow ({beta_slow})")
        self._factor = factor
        self._beta_fast = beta_fast
        self._beta_slow = beta_slow
        self._orig_max = original_max_position_embeddings

    def _dim_for_rotations(self, rotations: float, rope_base: int, head_dim: int) -> float:
        return head_dim * math.log(self._orig_max / (rotations * 2.0 * math.pi)) / (2.0 * math.log(rope_base))

    def inverse_frequencies(self, rope_base: int, head_dim: int) -> torch.Tensor:
        pairs = head_dim // 2
        plain = base_inverse_frequencies(rope_base, head_dim)
        lo = max(self._dim_for_rotations(self._beta_fast, rope_base, head_dim), 0.0)
        hi = min(self._dim_for_rotations(self._beta_slow, rope_base, head_dim), pairs - 1)
        ramp = ((torch.arange(pairs, dtype=torch.float32) - lo) / (hi - lo)).clamp_(0.0, 1.0)
        return plain + (plain / self._factor - plain) * ramp

    @property
    def attention_mscale(self) -> float:
        return 1.0 if self._factor <= 1.0 else 0.1 * math.log(self._factor) + 1.0
