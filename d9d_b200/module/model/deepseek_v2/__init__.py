from .decoder_layer import DeepseekV2Layer
from .huggingface import (
    DeepseekV2ExpertsFormat,
    mapper_from_huggingface_deepseek_v2,
    mapper_from_huggingface_deepseek_v2_for_causal_lm,
    mapper_from_huggingface_deepseek_v2_for_classification,
    mapper_from_huggingface_deepseek_v2_for_embedding,
    mapper_to_huggingface_deepseek_v2,
    mapper_to_huggingface_deepseek_v2_for_causal_lm,
    mapper_to_huggingface_deepseek_v2_for_classification,
    mapper_to_huggingface_deepseek_v2_for_embedding,
)
from .model import DeepseekV2ForCausalLM, DeepseekV2ForClassification, DeepseekV2ForEmbedding, DeepseekV2Model
from .params import (
    DeepseekV2ForCausalLMParameters,
    DeepseekV2ForClassificationParameters,
    DeepseekV2ForEmbeddingParameters,
    DeepseekV2LayerParameters,
    DeepseekV2Parameters,
)

__all__ = [
    "DeepseekV2ExpertsFormat",
    "DeepseekV2ForCausalLM",
    "DeepseekV2ForCausalLMParameters",
    "DeepseekV2ForClassification",
    "DeepseekV2ForClassificationParameters",
    "DeepseekV2ForEmbedding",
    "DeepseekV2ForEmbeddingParameters",
    "DeepseekV2Layer",
    "DeepseekV2LayerParameters",
    "DeepseekV2Model",
    "DeepseekV2Parameters",
    "mapper_from_huggingface_deepseek_v2",
    "mapper_from_huggingface_deepseek_v2_for_causal_lm",
    "mapper_from_huggingface_deepseek_v2_for_classification",
    "mapper_from_huggingface_deepseek_v2_for_embedding",
    "mapper_to_huggingface_deepseek_v2",
    "mapper_to_huggingface_deepseek_v2_for_causal_lm",
    "mapper_to_huggingface_deepseek_v2_for_classification",
    "mapper_to_huggingface_deepseek_v2_for_embedding",
]
