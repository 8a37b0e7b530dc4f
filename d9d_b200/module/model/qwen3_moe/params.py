"""Pydantic parameters of the Qwen3MoE family."""

from pydantic import BaseModel


class Qwen3MoELayerParameters(BaseModel):
    hidden_size: int
    intermediate_size: int
    num_experts: int
    experts_top_k: int
    num_attention_heads: int
    num_key_value_heads: int
    rms_norm_eps: float
    head_dim: int


class Qwen3MoEParameters(BaseModel):
    layer: Qwen3MoELayerParameters
    num_hidden_layers: int
    rope_base: int
    max_position_ids: int
    split_vocab_size: dict[str, int]
    split_vocab_order: list[str]
    pipeline_num_virtual_layers_pre: int = 0
    pipeline_num_virtual_layers_post: int = 0


class Qwen3MoEForCausalLMParameters(BaseModel):
    model: Qwen3MoEParameters


class Qwen3MoEForClassificationParameters(BaseModel):
    model: Qwen3MoEParameters
    num_labels: int
    classifier_dropout: float


class Qwen3MoEForEmbeddingParameters(BaseModel):
    model: Qwen3MoEParameters
    embedding_dim: int | None = None
    normalize: bool = False
