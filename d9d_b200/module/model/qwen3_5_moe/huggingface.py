"""HuggingFace ⇄ native state mappers of the Qwen3.5-MoE text family."""

from __future__ import annotations

from d9d_b200.model_state.mapper import ModelStateMapper
from d9d_b200.module.model import _huggingface as hf
from d9d_b200.module.model.qwen3_5.huggingface import _token_mixer

from .params import (
    Qwen3_5MoEForCausalLMParameters,
    Qwen3_5MoEForClassificationParameters,
    Qwen3_5MoEForEmbeddingParameters,
    Qwen3_5MoEParameters,
)

Qwen3_5MoEExpertsFormat = hf.ExpertsFormat


def _feed_forward(num_experts: int, experts_format: hf.ExpertsFormat) -> tuple[hf.Rule, ...]:
    if experts_format == hf.ExpertsFormat.MODULE_LIST:
        experts: hf.Rule = hf.ExpertsPerModule(
            hf_pattern="mlp.experts.{e}.{proj}.weight", native_pattern="mlp.grouped_experts.{proj}.weight",
            projections=(("gate_proj", "gate_proj"), ("up_proj", "up_proj"), ("down_proj", "down_proj")), num_experts=num_experts)
    elif experts_format == hf.ExpertsFormat.FUSED:
        experts = hf.ExpertsFused(hf_gate_up="mlp.experts.gate_up_proj", hf_down="mlp.experts.down_proj",
                                  native_gate="mlp.grouped_experts.gate_proj.weight", native_up="mlp.grouped_experts.up_proj.weight",
                                  native_down="mlp.grouped_experts.down_proj.weight")
    else:
        raise ValueError(f"Unsupported experts format {experts_format}")
    return (hf.Renamed("mlp.gate.weight", "mlp.router.gate.weight"), experts, *hf.shared_expert_rules("mlp.shared_expert"),
            hf.Renamed("mlp.shared_expert_gate.weight", "mlp.shared_expert.gate.weight"))


def _backbone(params: Qwen3_5MoEParameters, experts_format: hf.ExpertsFormat) -> tuple[hf.Rule, ...]:
    def layer_rules(index: int) -> tuple[hf.Rule, ...]:
        return (*_token_mixer(params.layer, index), *hf.norm_rules(), *_feed_forward(params.layer.num_experts, experts_format))

    return hf.backbone_rules(layer_rules, params.num_hidden_layers, hf.single_vocab_name(params.split_vocab_order))


def mapper_from_huggingface_qwen3_5_moe(params: Qwen3_5MoEParameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(_backbone(params, experts_format), hf.Direction.FROM_HF)


def mapper_to_huggingface_qwen3_5_moe(params: Qwen3_5MoEParameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(_backbone(params, experts_format), hf.Direction.TO_HF)


def _causal(params: Qwen3_5MoEForCausalLMParameters, experts_format: hf.ExpertsFormat) -> tuple[hf.Rule, ...]:
    return hf.causal_lm_rules(_backbone(params.model, experts_format), hf.single_vocab_name(params.model.split_vocab_order))


def mapper_from_huggingface_qwen3_5_moe_for_causal_lm(params: Qwen3_5MoEForCausalLMParameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(_causal(params, experts_format), hf.Direction.FROM_HF)


def mapper_to_huggingface_qwen3_5_moe_for_causal_lm(params: Qwen3_5MoEForCausalLMParameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(_causal(params, experts_format), hf.Direction.TO_HF)


def mapper_from_huggingface_qwen3_5_moe_for_classification(params: Qwen3_5MoEForClassificationParameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(hf.classification_rules(_backbone(params.model, experts_format)), hf.Direction.FROM_HF)


def mapper_to_huggingface_qwen3_5_moe_for_classification(params: Qwen3_5MoEForClassificationParameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(hf.classification_rules(_backbone(params.model, experts_format)), hf.Direction.TO_HF)


def mapper_from_huggingface_qwen3_5_moe_for_embedding(params: Qwen3_5MoEForEmbeddingParameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(hf.embedding_rules(_backbone(params.model, experts_format)), hf.Direction.FROM_HF)


def mapper_to_huggingface_qwen3_5_moe_for_embedding(params: Qwen3_5MoEForEmbeddingParameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(hf.embedding_rules(_backbone(params.model, experts_format)), hf.Direction.TO_HF)
