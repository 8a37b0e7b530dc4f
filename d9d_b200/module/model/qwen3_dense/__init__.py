from .decoder_layer import Qwen3DenseLayer
from .model import Qwen3DenseForCausalLM, Qwen3DenseForClassification, Qwen3DenseForEmbedding, Qwen3DenseModel
from .params import (
    Qwen3DenseForCausalLMParameters,
    Qwen3DenseForClassificationParameters,
    Qwen3DenseForEmbeddingParameters,
    Qwen3DenseLayerParameters,
    Qwen3DenseParameters,
)

__all__ = [
    "Qwen3DenseForCausalLM",
    "Qwen3DenseForCausalLMParameters",
    "Qwen3DenseForClassification",
    "Qwen3DenseForClassificationParameters",
    "Qwen3DenseForEmbedding",
    "Qwen3DenseForEmbeddingParameters",
    "Qwen3DenseLayer",
    "Qwen3DenseLayerParameters",
    "Qwen3DenseModel",
    "Qwen3DenseParameters",
]
