from .decoder_layer import MixtralLayer
from .model import MixtralForCausalLM, MixtralForClassification, MixtralForEmbedding, MixtralModel
from .params import (
    MixtralForCausalLMParameters,
    MixtralForClassificationParameters,
    MixtralForEmbeddingParameters,
    MixtralLayerParameters,
    MixtralParameters,
)

__all__ = [
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
]
