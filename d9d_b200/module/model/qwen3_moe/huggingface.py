"""HuggingFace ⇄ native state mappers of the Qwen3MoE family (rules in ``module/model/_huggingface.py``)."""

from __future__ import annotations

from d9d_b200.model_state.mapper import ModelStateMapper
from d9d_b200.module.model import _huggingface as hf

from .params import (
    Qwen3MoEForCausalLMParameters,
    Qwen3MoEForClassificationParameters,
    Qwen3MoEForEmbeddingParameters,
    Qwen3MoELayerParameters,
    Qwen3MoEParameters,
)

Qwen3MoEExpertsFormat = hf.ExpertsFormat


def _experts(layer: Qwen3MoELayerParameters, experts_format: hf.ExpertsFormat) -> tuple[hf.Rule, ...]:
    if experts_format == hf.ExpertsFormat.MODULE_LIST:
        return (hf.ExpertsPerModule(hf_pattern="mlp.experts.{e}.{proj}.weight",
                                    projections=(("gate_proj", "gate_proj"), ("up_proj", "up_proj"), ("down_proj", "down_proj")),
                                    native_pattern="mlp.grouped_experts.{proj}.weight", num_experts=layer.num_experts),)
    if experts_format == hf.ExpertsFormat.FUSED:
        return (hf.ExpertsFused(hf_gate_up="mlp.experts.gate_up_proj", hf_down="mlp.experts.down_proj",
                                native_gate="mlp.grouped_experts.gate_proj.weight", native_up="mlp.grouped_experts.up_proj.weight",
                                native_down="mlp.grouped_experts.down_proj.weight"),)
    raise ValueError(f"Unsupported experts format {experts_format}")


def _backbone(params: Qwen3MoEParameters, experts_format: hf.ExpertsFormat) -> tuple[hf.Rule, ...]:
    layer = (*hf.attention_rules(qk_norm=True), *hf.norm_rules(), hf.Renamed("mlp.gate.weight", "mlp.router.gate.weight"),
             *_experts(params.layer, experts_format))
    return hf.backbone_rules(layer, params.num_hidden_layers, hf.single_vocab_name(params.split_vocab_order))


def mapper_from_huggingface_qwen3_moe(params: Qwen3MoEParameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(_backbone(params, experts_format), hf.Direction.FROM_HF)


def mapper_to_huggingface_qwen3_moe(params: Qwen3MoEParameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(_backbone(params, experts_format), hf.Direction.TO_HF)


def _causal(params: Qwen3MoEForCausalLMParameters, experts_format: hf.ExpertsFormat) -> tuple[hf.Rule, ...]:
    return hf.causal_lm_rules(_backbone(params.model, experts_format), hf.single_vocab_name(params.model.split_vocab_order))


def mapper_from_huggingface_qwen3_moe_for_causal_lm(params: Qwen3MoEForCausalLMParameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(_causal(params, experts_format), hf.Direction.FROM_HF)


def mapper_to_huggingface_qwen3_moe_for_causal_lm(params: Qwen3MoEForCausalLMParameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(_causal(params, experts_format), hf.Direction.TO_HF)


def mapper_from_huggingface_qwen3_moe_for_classification(params: Qwen3MoEForClassificationParameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(hf.classification_rules(_backbone(params.model, experts_format)), hf.Direction.FROM_HF)


def mapper_to_huggingface_qwen3_moe_for_classification(params: Qwen3MoEForClassificationParameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(hf.classification_rules(_backbone(params.model, experts_format)), hf.Direction.TO_HF)


def mapper_from_huggingface_qwen3_moe_for_embedding(params: Qwen3MoEForEmbeddingParameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(hf.embedding_rules(_backbone(params.model, experts_format)), hf.Direction.FROM_HF)


def mapper_to_huggingface_qwen3_moe_for_embedding(params: Qwen3MoEForEmbeddingParameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(hf.embedding_rules(_backbone(params.model, experts_format)), hf.Direction.TO_HF)
