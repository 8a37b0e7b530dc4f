"""Hyper-parameters of the Qwen3.5 (dense, text) family: a hybrid stack in which every ``full_attention_interval``-th layer
uses gated softmax attention with partial rotary embedding and the others Gated DeltaNet linear attention."""

from pydantic import BaseModel, Field, PositiveInt, model_validator

from d9d_b200.module.model._params import family_parameters


class HybridMixerFields(BaseModel):
    """Fields describing the token mixers of a hybrid layer stack (shared by the dense and the MoE variant)."""

    hidden_size: PositiveInt
    rms_norm_eps: float
    # full (softmax) attention layers
    num_attention_heads: PositiveInt
    num_key_value_heads: PositiveInt
    head_dim: PositiveInt
    partial_rotary_factor: float = Field(default=0.25, gt=0.0, le=1.0)
    # linear attention layers
    linear_num_key_heads: PositiveInt
    linear_num_value_heads: PositiveInt
    linear_key_head_dim: PositiveInt
    linear_value_head_dim: PositiveInt
    linear_conv_kernel_dim: PositiveInt = 4
    # layer ``i`` (0-based) uses full attention iff ``(i + 1) % full_attention_interval == 0``
    full_attention_interval: PositiveInt = 4

    @model_validator(mode="after")
    def _check(self):  # noqa: ANN202
        if self.num_attention_heads % self.num_key_value_heads != 0:
            raise ValueError("num_attention_heads must be a multiple of num_key_value_heads")
        if self.linear_num_value_heads % self.linear_num_key_heads != 0:
            raise ValueError("linear_num_value_heads must be a multiple of linear_num_key_heads")
        if self.rope_dim % 2 != 0:
            raise ValueError("head_dim * partial_rotary_factor must be even")
        return self

    @property
    def rope_dim(self) -> int:
        return int(self.head_dim * self.partial_rotary_factor)

    def uses_full_attention(self, layer_index: int) -> bool:
        return (layer_index + 1) % self.full_attention_interval == 0


class HybridLayerFields(HybridMixerFields):
    intermediate_size: PositiveInt


_generated = family_parameters("Qwen3_5", HybridLayerFields, __name__)

Qwen3_5LayerParameters = _generated["Qwen3_5LayerParameters"]
Qwen3_5Parameters = _generated["Qwen3_5Parameters"]
Qwen3_5ForCausalLMParameters = _generated["Qwen3_5ForCausalLMParameters"]
Qwen3_5ForClassificationParameters = _generated["Qwen3_5ForClassificationParameters"]
Qwen3_5ForEmbeddingParameters = _generated["Qwen3_5ForEmbeddingParameters"]

__all__ = list(_generated)
