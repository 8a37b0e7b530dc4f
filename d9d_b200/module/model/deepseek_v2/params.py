"""Hyper-parameters of the DeepSeek-V2 / V3 family: multi-head latent attention, dense first layers, then MoE layers with
always-on shared experts; ``router`` selects V3's sigmoid / group-limited / bias-corrected routing (classes generated from
the field set below by ``module/model/_params.py``)."""

from pydantic import BaseModel, NonNegativeInt, PositiveInt, model_validator

from d9d_b200.module.block.moe.router import RouterParameters
from d9d_b200.module.model._params import family_parameters


class LatentMoELayerFields(BaseModel):
    hidden_size: PositiveInt
    rms_norm_eps: float
    # multi-head latent attention
    num_attention_heads: PositiveInt
    qk_nope_head_dim: PositiveInt
    qk_rope_head_dim: PositiveInt
    v_head_dim: PositiveInt
    kv_lora_rank: PositiveInt
    q_lora_rank: PositiveInt | None
    # feed-forward: layers ``< first_k_dense_replace`` are dense SwiGLU MLPs, the rest MoE
    intermediate_size: PositiveInt
    first_k_dense_replace: NonNegativeInt
    moe_intermediate_size: PositiveInt  # per routed expert; shared experts are ``num_shared_experts`` times as wide
    num_experts: PositiveInt
    experts_top_k: PositiveInt
    num_shared_experts: NonNegativeInt
    router_renormalize_probabilities: bool = False
    # DeepSeek-V3 routing: sigmoid scores, selection bias, group-limited top-k, weight scaling (defaults = DeepSeek-V2 greedy)
    router: RouterParameters = RouterParameters()

    @model_validator(mode="after")
    def _check(self):  # noqa: ANN202
        if self.experts_top_k > self.num_experts:
            raise ValueError("experts_top_k cannot exceed num_experts")
        if self.v_head_dim > self.qk_nope_head_dim + self.qk_rope_head_dim:
            raise ValueError("v_head_dim cannot exceed the query/key head dim")
        return self

    @property
    def head_dim(self) -> int:
        """Width of the rotary tables: only the decoupled rope part of a head rotates."""
        return self.qk_rope_head_dim


_generated = family_parameters("DeepseekV2", LatentMoELayerFields, __name__)

DeepseekV2LayerParameters = _generated["DeepseekV2LayerParameters"]
DeepseekV2Parameters = _generated["DeepseekV2Parameters"]
DeepseekV2ForCausalLMParameters = _generated["DeepseekV2ForCausalLMParameters"]
DeepseekV2ForClassificationParameters = _generated["DeepseekV2ForClassificationParameters"]
DeepseekV2ForEmbeddingParameters = _generated["DeepseekV2ForEmbeddingParameters"]

__all__ = list(_generated)
