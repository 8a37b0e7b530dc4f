from .huggingface import (
    Qwen3_5MoEExpertsFormat,
    mapper_from_huggingface_qwen3_5_moe,
    mapper_from_huggingface_qwen3_5_moe_for_causal_lm,
    mapper_from_huggingface_qwen3_5_moe_for_classification,
    mapper_from_huggingface_qwen3_5_moe_for_embedding,
    mapper_to_huggingface_qwen3_5_moe,
    mapper_to_huggingface_qwen3_5_moe_for_causal_lm,
    mapper_to_huggingface_qwen3_5_moe_for_classification,
    mapper_to_huggingface_qwen3_5_moe_for_embedding,
)
from .model import Qwen3_5MoEForCausalLM, Qwen3_5MoEForClassification, Qwen3_5MoEForEmbedding, Qwen3_5MoELayer, Qwen3_5MoEModel
from .params import (
    Qwen3_5MoEForCausalLMParameters,
    Qwen3_5MoEForClassificationParameters,
    Qwen3_5MoEForEmbeddingParameters,
    Qwen3_5MoELayerParameters,
    Qwen3_5MoEParameters,
)

__all__ = [
    "Qwen3_5MoEExpertsFormat",
    "Qwen3_5MoEForCausalLM",
    "Qwen3_5MoEForCausalLMParameters",
    "Qwen3_5MoEForClassification",
    "Qwen3_5MoEForClassificationParameters",
    "Qwen3_5MoEForEmbedding",
    "Qwen3_5MoEForEmbeddingParameters",
    "Qwen3_5MoELayer",
    "Qwen3_5MoELayerParameters",
    "Qwen3_5MoEModel",
    "Qwen3_5MoEParameters",
    "mapper_from_huggingface_qwen3_5_moe",
    "mapper_from_huggingface_qwen3_5_moe_for_causal_lm",
    "mapper_from_huggingface_qwen3_5_moe_for_classification",
    "mapper_from_huggingface_qwen3_5_moe_for_embedding",
    "mapper_to_huggingface_qwen3_5_moe",
    "mapper_to_huggingface_qwen3_5_moe_for_causal_lm",
    "mapper_to_huggingface_qwen3_5_moe_for_classification",
    "mapper_to_huggingface_qwen3_5_moe_for_embedding",
]
