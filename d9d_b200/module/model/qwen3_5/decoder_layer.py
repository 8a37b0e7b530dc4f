from __future__ import annotations

from d9d_b200.module.block.attention import GatedDeltaNet, GroupedQueryAttention
from d9d_b200.module.block.attention.linear.gated_deltanet import MambaDecayGateParameters
from d9d_b200.module.block.ffn import SwiGLU
from d9d_b200.module.block.positional import RotaryEmbeddingStyle
from d9d_b200.module.model.decoder import PreNormDecoderLayer

from .params import Qwen3_5LayerParameters


def build_token_mixer(params, index: int):  # noqa: ANN001, ANN201
    """Gated DeltaNet, or - every ``full_attention_interval``-th layer - gated softmax attention with zero-centred q/k norms
    and partial rotary embedding."""
    if params.uses_full_attention(index):
        return GroupedQueryAttention(
            hidden_size=params.hidden_size, num_attention_heads=params.num_attention_heads,
            num_key_value_heads=params.num_key_value_heads, head_dim=params.head_dim, qk_norm_eps=params.rms_norm_eps,
            is_causal=True, rope_style=RotaryEmbeddingStyle.HALF, rope_dim=params.rope_dim, enable_output_gate=True,
            qk_norm_zero_centered=True)
    return GatedDeltaNet(
        hidden_size=params.hidden_size, num_query_key_heads=params.linear_num_key_heads,
        num_value_heads=params.linear_num_value_heads, head_qk_dim=params.linear_key_head_dim,
        head_v_dim=params.linear_value_head_dim, norm_eps=params.rms_norm_eps, conv_size=params.linear_conv_kernel_dim,
        decay_gate=MambaDecayGateParameters(normalizer=16.0, dt_min=0.001, dt_max=0.1, dt_init_floor=1e-4))


class Qwen3_5Layer(PreNormDecoderLayer):
    """Hybrid decoder layer: Gated DeltaNet linear attention, or - every ``full_attention_interval``-th layer - gated
    softmax attention with zero-centred q/k norms and partial rotary embedding; zero-centred RMSNorms; dense SwiGLU MLP."""

    def __init__(self, params: Qwen3_5LayerParameters, index: int):
        super().__init__(build_token_mixer(params, index), SwiGLU(params.hidden_size, params.intermediate_size), params.hidden_size,
                         params.rms_norm_eps, zero_centered_norm=True)
