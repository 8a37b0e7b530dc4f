"""Pydantic parameters of the Llama3 family."""

from pydantic import BaseModel


class Llama3LayerParameters(BaseModel):
    hidden_size: int
    intermediate_size: int
    num_attention_heads: int
    num_key_value_heads: int
    rms_norm_eps: float
    head_dim: int


class Llama3Parameters(BaseModel):
    layer: Llama3LayerParameters
    num_hidden_layers: int
    rope_base: int
    max_position_ids: int
    split_vocab_size: dict[str, int]
    split_vocab_order: list[str]
    pipeline_num_virtual_layers_pre: int = 0
    pipeline_num_virtual_layers_post: int = 0


class Llama3ForCausalLMParameters(BaseModel):
    model: Llama3Parameters


class Llama3ForClassificationParameters(BaseModel):
    model: Llama3Parameters
    num_labels: int
    classifier_dropout: float


class Llama3ForEmbeddingParameters(BaseModel):
    model: Llama3Parameters
    embedding_dim: int | None = None
    normalize: bool = False
