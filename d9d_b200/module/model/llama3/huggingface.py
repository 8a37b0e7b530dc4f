"""HuggingFace ⇄ native state mappers of the Llama3 family (rules in ``module/model/_huggingface.py``)."""

from __future__ import annotations

from d9d_b200.model_state.mapper import ModelStateMapper
from d9d_b200.module.model import _huggingface as hf

from .params import Llama3ForCausalLMParameters, Llama3ForClassificationParameters, Llama3ForEmbeddingParameters, Llama3Parameters


def _backbone(params: Llama3Parameters) -> tuple[hf.Rule, ...]:
    layer = (*hf.attention_rules(qk_norm=False), *hf.norm_rules(), *hf.dense_mlp_rules())
    return hf.backbone_rules(layer, params.num_hidden_layers, hf.single_vocab_name(params.split_vocab_order))


def mapper_from_huggingface_llama3(params: Llama3Parameters) -> ModelStateMapper:
    return hf.compile_rules(_backbone(params), hf.Direction.FROM_HF)


def mapper_to_huggingface_llama3(params: Llama3Parameters) -> ModelStateMapper:
    return hf.compile_rules(_backbone(params), hf.Direction.TO_HF)


def mapper_from_huggingface_llama3_for_causal_lm(params: Llama3ForCausalLMParameters) -> ModelStateMapper:
    return hf.compile_rules(hf.causal_lm_rules(_backbone(params.model), hf.single_vocab_name(params.model.split_vocab_order)), hf.Direction.FROM_HF)


def mapper_to_huggingface_llama3_for_causal_lm(params: Llama3ForCausalLMParameters) -> ModelStateMapper:
    return hf.compile_rules(hf.causal_lm_rules(_backbone(params.model), hf.single_vocab_name(params.model.split_vocab_order)), hf.Direction.TO_HF)


def mapper_from_huggingface_llama3_for_classification(params: Llama3ForClassificationParameters) -> ModelStateMapper:
    return hf.compile_rules(hf.classification_rules(_backbone(params.model)), hf.Direction.FROM_HF)


def mapper_to_huggingface_llama3_for_classification(params: Llama3ForClassificationParameters) -> ModelStateMapper:
    return hf.compile_rules(hf.classification_rules(_backbone(params.model)), hf.Direction.TO_HF)


def mapper_from_huggingface_llama3_for_embedding(params: Llama3ForEmbeddingParameters) -> ModelStateMapper:
    return hf.compile_rules(hf.embedding_rules(_backbone(params.model)), hf.Direction.FROM_HF)


def mapper_to_huggingface_llama3_for_embedding(params: Llama3ForEmbeddingParameters) -> ModelStateMapper:
    return hf.compile_rules(hf.embedding_rules(_backbone(params.model)), hf.Direction.TO_HF)
