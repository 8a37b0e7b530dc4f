from __future__ import annotations

from d9d_b200.module.block.attention import GroupedQueryAttention
from d9d_b200.module.block.ffn import SwiGLU
from d9d_b200.module.block.positional import RotaryEmbeddingStyle
from d9d_b200.module.model.decoder import PreNormDecoderLayer

from .params import Qwen3DenseLayerParameters


class Qwen3DenseLayer(PreNormDecoderLayer):
    """One pre-norm decoder layer: causal GQA (with per-head q/k RMSNorm) + SwiGLU MLP."""

    def __init__(self, params: Qwen3DenseLayerParameters):
        attn = GroupedQueryAttention(
            hidden_size=params.hidden_size,
            num_attention_heads=params.num_attention_heads,
            num_key_value_heads=params.num_key_value_heads,
            head_dim=params.head_dim,
            qk_norm_eps=params.rms_norm_eps,
            is_causal=True,
            rope_style=RotaryEmbeddingStyle.HALF,
        )
        mlp = SwiGLU(hidden_size=params.hidden_size, intermediate_size=params.intermediate_size, bias=False)
        super().__init__(attn, mlp, params.hidden_size, params.rms_norm_eps)
