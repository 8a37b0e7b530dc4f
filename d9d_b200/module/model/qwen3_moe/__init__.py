from .decoder_layer import Qwen3MoELayer
from .model import Qwen3MoEForCausalLM, Qwen3MoEForClassification, Qwen3MoEForEmbedding, Qwen3MoEModel
from .params import (
    Qwen3MoEForCausalLMParameters,
    Qwen3MoEForClassificationParameters,
    Qwen3MoEForEmbeddingParameters,
    Qwen3MoELayerParameters,
    Qwen3MoEParameters,
)

__all__ = [
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
]
