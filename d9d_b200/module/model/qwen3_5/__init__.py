from .decoder_layer import Qwen3_5Layer
from .huggingface import (
    mapper_from_huggingface_qwen3_5,
    mapper_from_huggingface_qwen3_5_for_causal_lm,
    mapper_from_huggingface_qwen3_5_for_classification,
    mapper_from_huggingface_qwen3_5_for_embedding,
    mapper_to_huggingface_qwen3_5,
    mapper_to_huggingface_qwen3_5_for_causal_lm,
    mapper_to_huggingface_qwen3_5_for_classification,
    mapper_to_huggingface_qwen3_5_for_embedding,
)
from .model import Qwen3_5ForCausalLM, Qwen3_5ForClassification, Qwen3_5ForEmbedding, Qwen3_5Model
from .params import (
    Qwen3_5ForCausalLMParameters,
    Qwen3_5ForClassificationParameters,
    Qwen3_5ForEmbeddingParameters,
    Qwen3_5LayerParameters,
    Qwen3_5Parameters,
)

__all__ = [
    "Qwen3_5ForCausalLM",
    "Qwen3_5ForCausalLMParameters",
    "Qwen3_5ForClassification",
    "Qwen3_5ForClassificationParameters",
    "Qwen3_5ForEmbedding",
    "Qwen3_5ForEmbeddingParameters",
    "Qwen3_5Layer",
    "Qwen3_5LayerParameters",
    "Qwen3_5Model",
    "Qwen3_5Parameters",
    "mapper_from_huggingface_qwen3_5",
    "mapper_from_huggingface_qwen3_5_for_causal_lm",
    "mapper_from_huggingface_qwen3_5_for_classification",
    "mapper_from_huggingface_qwen3_5_for_embedding",
    "mapper_to_huggingface_qwen3_5",
    "mapper_to_huggingface_qwen3_5_for_causal_lm",
    "mapper_to_huggingface_qwen3_5_for_classification",
    "mapper_to_huggingface_qwen3_5_for_embedding",
]
