"""Pydantic parameters of the Mixtral family."""

from pydantic import BaseModel


class MixtralLayerParameters(BaseModel):
    hidden_size: int
    intermediate_size: int
    num_experts: int
    experts_top_k: int
    num_attention_heads: int
    num_key_value_heads: int
    rms_norm_eps: float
    head_dim: int


class MixtralParameters(BaseModel):
    layer: MixtralLayerParameters
    num_hidden_layers: int
    rope_base: int
    max_position_ids: int
    split_vocab_size: dict[str, int]
    split_vocab_order: list[str]
    pipeline_num_virtual_layers_pre: int = 0
    pipeline_num_virtual_layers_post: int = 0


class MixtralForCausalLMParameters(BaseModel):
    model: MixtralParameters


class MixtralForClassificationParameters(BaseModel):
    model: MixtralParameters
    num_labels: int
    classifier_dropout: float


class MixtralForEmbeddingParameters(BaseModel):
    model: MixtralParameters
    embedding_dim: int | None = None
    normalize: bool = False
