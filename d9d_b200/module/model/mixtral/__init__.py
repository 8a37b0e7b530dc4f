from .decoder_layer import MixtralLayer
from .huggingface import (
    MixtralExpertsFormat,
    mapper_from_huggingface_mixtral,
    mapper_from_huggingface_mixtral_for_causal_lm,
    mapper_from_huggingface_mixtral_for_classification,
    mapper_from_huggingface_mixtral_for_embedding,
    mapper_to_huggingface_mixtral,
    mapper_to_huggingface_mixtral_for_causal_lm,
    mapper_to_huggingface_mixtral_for_classification,
    mapper_to_huggingface_mixtral_for_embedding,
)
from .model import MixtralForCausalLM, MixtralForClassification, MixtralForEmbedding, MixtralModel
from .params import (
    MixtralForCausalLMParameters,
    MixtralForClassificationParameters,
    MixtralForEmbeddingParameters,
    MixtralLayerParameters,
    MixtralParameters,
)

__all__ = [
    "MixtralExpertsFormat",
    "MixtralForCausalLM",
    "MixtralForCausalLMParameters",
    "MixtralForClassification",
    "MixtralForClassificationParameters",
    "MixtralForEmbedding",
    "MixtralForEmbeddingParameters",
    "MixtralLayer",
    "MixtralLayerParameters",
    "MixtralModel",
    "MixtralParameters",
    "mapper_from_huggingface_mixtral",
    "mapper_from_huggingface_mixtral_for_causal_lm",
    "mapper_from_huggingface_mixtral_for_classification",
    "mapper_from_huggingface_mixtral_for_embedding",
    "mapper_to_huggingface_mixtral",
    "mapper_to_huggingface_mixtral_for_causal_lm",
    "mapper_to_huggingface_mixtral_for_classification",
    "mapper_to_huggingface_mixtral_for_embedding",
]
