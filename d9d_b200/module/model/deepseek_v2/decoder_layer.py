from __future__ import annotations

from d9d_b200.module.block.attention import MultiHeadLatentAttention
from d9d_b200.module.block.ffn import SwiGLU
from d9d_b200.module.block.moe import MoELayer, SharedExpertParameters
from d9d_b200.module.block.positional import RotaryEmbeddingStyle
from d9d_b200.module.model.decoder import PreNormDecoderLayer

from .params import DeepseekV2LayerParameters


class DeepseekV2Layer(PreNormDecoderLayer):
    """Pre-norm decoder layer with multi-head latent attention; layer ``index`` decides between the dense MLP of the first
    ``first_k_dense_replace`` layers and the MoE block (softmax router, greedy top-k, shared experts) of the rest."""

    def __init__(self, params: DeepseekV2LayerParameters, index: int):
        attention = MultiHeadLatentAttention(
            hidden_size=params.hidden_size, num_attention_heads=params.num_attention_heads,
            qk_nope_head_dim=params.qk_nope_head_dim, qk_rope_head_dim=params.qk_rope_head_dim, v_head_dim=params.v_head_dim,
            kv_lora_rank=params.kv_lora_rank, q_lora_rank=params.q_lora_rank, qk_down_norm_eps=params.rms_norm_eps,
            is_causal=True, rope_style=RotaryEmbeddingStyle.INTERLEAVED)
        if index < params.first_k_dense_replace:
            mlp = SwiGLU(params.hidden_size, params.intermediate_size)
        else:
            shared = (SharedExpertParameters(intermediate_size=params.moe_intermediate_size * params.num_shared_experts, enable_gate=False)
                      if params.num_shared_experts > 0 else None)
            mlp = MoELayer(hidden_dim=params.hidden_size, intermediate_dim_grouped=params.moe_intermediate_size,
                           num_grouped_experts=params.num_experts, top_k=params.experts_top_k,
                           router_renormalize_probabilities=params.router_renormalize_probabilities, shared_expert=shared,
                           router=params.router)
        super().__init__(attention, mlp, params.hidden_size, params.rms_norm_eps)
