"""HuggingFace ⇄ native state mappers of the Qwen3Dense family (rules in ``module/model/_huggingface.py``)."""

from __future__ import annotations

from d9d_b200.model_state.mapper import ModelStateMapper
from d9d_b200.module.model import _huggingface as hf

from .params import Qwen3DenseForCausalLMParameters, Qwen3DenseForClassificationParameters, Qwen3DenseForEmbeddingParameters, Qwen3DenseParameters


def _backbone(params: Qwen3DenseParameters) -> tuple[hf.Rule, ...]:
    layer = (*hf.attention_rules(qk_norm=True), *hf.norm_rules(), *hf.dense_mlp_rules())
    return hf.backbone_rules(layer, params.num_hidden_layers, hf.single_vocab_name(params.split_vocab_order))


def mapper_from_huggingface_qwen3_dense(params: Qwen3DenseParameters) -> ModelStateMapper:
    return hf.compile_rules(_backbone(params), hf.Direction.FROM_HF)


def mapper_to_huggingface_qwen3_dense(params: Qwen3DenseParameters) -> ModelStateMapper:
    return hf.compile_rules(_backbone(params), hf.Direction.TO_HF)


def mapper_from_huggingface_qwen3_dense_for_causal_lm(params: Qwen3DenseForCausalLMParameters) -> ModelStateMapper:
    return hf.compile_rules(hf.causal_lm_rules(_backbone(params.model), hf.single_vocab_name(params.model.split_vocab_order)), hf.Direction.FROM_HF)


def mapper_to_huggingface_qwen3_dense_for_causal_lm(params: Qwen3DenseForCausalLMParameters) -> ModelStateMapper:
    return hf.compile_rules(hf.causal_lm_rules(_backbone(params.model), hf.single_vocab_name(params.model.split_vocab_order)), hf.Direction.TO_HF)


def mapper_from_huggingface_qwen3_dense_for_classification(params: Qwen3DenseForClassificationParameters) -> ModelStateMapper:
    return hf.compile_rules(hf.classification_rules(_backbone(params.model)), hf.Direction.FROM_HF)


def mapper_to_huggingface_qwen3_dense_for_classification(params: Qwen3DenseForClassificationParameters) -> ModelStateMapper:
    return hf.compile_rules(hf.classification_rules(_backbone(params.model)), hf.Direction.TO_HF)


def mapper_from_huggingface_qwen3_dense_for_embedding(params: Qwen3DenseForEmbeddingParameters) -> ModelStateMapper:
    return hf.compile_rules(hf.embedding_rules(_backbone(params.model)), hf.Direction.FROM_HF)


def mapper_to_huggingface_qwen3_dense_for_embedding(params: Qwen3DenseForEmbeddingParameters) -> ModelStateMapper:
    return hf.compile_rules(hf.embedding_rules(_backbone(params.model)), hf.Direction.TO_HF)
