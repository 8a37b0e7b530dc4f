"""Pydantic parameters of the Qwen3Dense family."""

from pydantic import BaseModel


class Qwen3DenseLayerParameters(BaseModel):
    hidden_size: int
    intermediate_size: int
    num_attention_heads: int
    num_key_value_heads: int
    rms_norm_eps: float
    head_dim: int


class Qwen3DenseParameters(BaseModel):
    layer: Qwen3DenseLayerParameters
    num_hidden_layers: int
    rope_base: int
    max_position_ids: int
    split_vocab_size: dict[str, int]
    split_vocab_order: list[str]
    pipeline_num_virtual_layers_pre: int = 0
    pipeline_num_virtual_layers_post: int = 0


class Qwen3DenseForCausalLMParameters(BaseModel):
    model: Qwen3DenseParameters


class Qwen3DenseForClassificationParameters(BaseModel):
    model: Qwen3DenseParameters
    num_labels: int
    classifier_dropout: float


class Qwen3DenseForEmbeddingParameters(BaseModel):
    model: Qwen3DenseParameters
    embedding_dim: int | None = None
    normalize: bool = False
