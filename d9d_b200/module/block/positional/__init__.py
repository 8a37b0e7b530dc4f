"""Rotary position embeddings and context-extension scalings."""

from .rope import RotaryEmbeddingApplicator, RotaryEmbeddingProvider, RotaryEmbeddingStyle, prepare_rotary_cos_sin_emb
from .rope_scaling import LinearRopeScaling, NoRopeScaling, NtkRopeScaling, RopeScaling, YarnRopeScaling

__all__ = [
    "LinearRopeScaling",
    "NoRopeScaling",
    "NtkRopeScaling",
    "RopeScaling",
    "RotaryEmbeddingApplicator",
    "RotaryEmbeddingProvider",
    "RotaryEmbeddingStyle",
    "YarnRopeScaling",
    "prepare_rotary_cos_sin_emb",
]
