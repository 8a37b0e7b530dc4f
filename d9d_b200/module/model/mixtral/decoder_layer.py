from __future__ import annotations

from d9d_b200.module.block.attention import GroupedQueryAttention
from d9d_b200.module.block.moe import MoELayer
from d9d_b200.module.block.positional import RotaryEmbeddingStyle
from d9d_b200.module.model.decoder import PreNormDecoderLayer

from .params import MixtralLayerParameters


class MixtralLayer(PreNormDecoderLayer):
    """One pre-norm decoder layer: causal GQA (without per-head q/k RMSNorm) + MoE MLP."""

    def __init__(self, params: MixtralLayerParameters):
        attn = GroupedQueryAttention(
            hidden_size=params.hidden_size,
            num_attention_heads=params.num_attention_heads,
            num_key_value_heads=params.num_key_value_heads,
            head_dim=params.head_dim,
            qk_norm_eps=None,
            is_causal=True,
            rope_style=RotaryEmbeddingStyle.HALF,
        )
        mlp = MoELayer(hidden_dim=params.hidden_size, num_grouped_experts=params.num_experts,
                         intermediate_dim_grouped=params.intermediate_size, top_k=params.experts_top_k,
                         router_renormalize_probabilities=True)
        super().__init__(attn, mlp, params.hidden_size, params.rms_norm_eps)
