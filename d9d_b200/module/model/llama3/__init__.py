from .decoder_layer import Llama3Layer
from .huggingface import (
    mapper_from_huggingface_llama3,
    mapper_from_huggingface_llama3_for_causal_lm,
    mapper_from_huggingface_llama3_for_classification,
    mapper_from_huggingface_llama3_for_embedding,
    mapper_to_huggingface_llama3,
    mapper_to_huggingface_llama3_for_causal_lm,
    mapper_to_huggingface_llama3_for_classification,
    mapper_to_huggingface_llama3_for_embedding,
)
from .model import Llama3ForCausalLM, Llama3ForClassification, Llama3ForEmbedding, Llama3Model
from .params import (
    Llama3ForCausalLMParameters,
    Llama3ForClassificationParameters,
    Llama3ForEmbeddingParameters,
    Llama3LayerParameters,
    Llama3Parameters,
)

__all__ = [
    "Llama3ForCausalLM",
    "Llama3ForCausalLMParameters",
    "Llama3ForClassification",
    "Llama3ForClassificationParameters",
    "Llama3ForEmbedding",
    "Llama3ForEmbeddingParameters",
    "Llama3Layer",
    "Llama3LayerParameters",
    "Llama3Model",
    "Llama3Parameters",
    "mapper_from_huggingface_llama3",
    "mapper_from_huggingface_llama3_for_causal_lm",
    "mapper_from_huggingface_llama3_for_classification",
    "mapper_from_huggingface_llama3_for_embedding",
    "mapper_to_huggingface_llama3",
    "mapper_to_huggingface_llama3_for_causal_lm",
    "mapper_to_huggingface_llama3_for_classification",
    "mapper_to_huggingface_llama3_for_embedding",
]
