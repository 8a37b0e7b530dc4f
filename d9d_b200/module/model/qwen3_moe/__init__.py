from .decoder_layer import Qwen3MoELayer
from .huggingface import (
    Qwen3MoEExpertsFormat,
    mapper_from_huggingface_qwen3_moe,
    mapper_from_huggingface_qwen3_moe_for_causal_lm,
    mapper_from_huggingface_qwen3_moe_for_classification,
    mapper_from_huggingface_qwen3_moe_for_embedding,
    mapper_to_huggingface_qwen3_moe,
    mapper_to_huggingface_qwen3_moe_for_causal_lm,
    mapper_to_huggingface_qwen3_moe_for_classification,
    mapper_to_huggingface_qwen3_moe_for_embedding,
)
from .model import Qwen3MoEForCausalLM, Qwen3MoEForClassification, Qwen3MoEForEmbedding, Qwen3MoEModel
from .params import (
    Qwen3MoEForCausalLMParameters,
    Qwen3MoEForClassificationParameters,
    Qwen3MoEForEmbeddingParameters,
    Qwen3MoELayerParameters,
    Qwen3MoEParameters,
)

__all__ = [
    "Qwen3MoEExpertsFormat",
    "Qwen3MoEForCausalLM",
    "Qwen3MoEForCausalLMParameters",
    "Qwen3MoEForClassification",
    "Qwen3MoEForClassificationParameters",
    "Qwen3MoEForEmbedding",
    "Qwen3MoEForEmbeddingParameters",
    "Qwen3MoELayer",
    "Qwen3MoELayerParameters",
    "Qwen3MoEModel",
    "Qwen3MoEParameters",
    "mapper_from_huggingface_qwen3_moe",
    "mapper_from_huggingface_qwen3_moe_for_causal_lm",
    "mapper_from_huggingface_qwen3_moe_for_classification",
    "mapper_from_huggingface_qwen3_moe_for_embedding",
    "mapper_to_huggingface_qwen3_moe",
    "mapper_to_huggingface_qwen3_moe_for_causal_lm",
    "mapper_to_huggingface_qwen3_moe_for_classification",
    "mapper_to_huggingface_qwen3_moe_for_embedding",
]
