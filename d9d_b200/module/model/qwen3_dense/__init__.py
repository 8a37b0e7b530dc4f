from .decoder_layer import Qwen3DenseLayer
from .huggingface import (
    mapper_from_huggingface_qwen3_dense,
    mapper_from_huggingface_qwen3_dense_for_causal_lm,
    mapper_from_huggingface_qwen3_dense_for_classification,
    mapper_from_huggingface_qwen3_dense_for_embedding,
    mapper_to_huggingface_qwen3_dense,
    mapper_to_huggingface_qwen3_dense_for_causal_lm,
    mapper_to_huggingface_qwen3_dense_for_classification,
    mapper_to_huggingface_qwen3_dense_for_embedding,
)
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
    "mapper_from_huggingface_qwen3_dense",
    "mapper_from_huggingface_qwen3_dense_for_causal_lm",
    "mapper_from_huggingface_qwen3_dense_for_classification",
    "mapper_from_huggingface_qwen3_dense_for_embedding",
    "mapper_to_huggingface_qwen3_dense",
    "mapper_to_huggingface_qwen3_dense_for_causal_lm",
    "mapper_to_huggingface_qwen3_dense_for_classification",
    "mapper_to_huggingface_qwen3_dense_for_embedding",
]
