"""HuggingFace ⇄ native state mappers of the DeepSeek-V2 family (rules in ``module/model/_huggingface.py``)."""

from __future__ import annotations

from d9d_b200.model_state.mapper import ModelStateMapper
from d9d_b200.module.model import _huggingface as hf

from .params import (
    DeepseekV2ForCausalLMParameters,
    DeepseekV2ForClassificationParameters,
    DeepseekV2ForEmbeddingParameters,
    DeepseekV2LayerParameters,
    DeepseekV2Parameters,
)

DeepseekV2ExpertsFormat = hf.ExpertsFormat


def _feed_forward(layer: DeepseekV2LayerParameters, index: int, experts_format: hf.ExpertsFormat) -> tuple[hf.Rule, ...]:
    if index < layer.first_k_dense_replace:
        return hf.dense_mlp_rules()
    if experts_format == hf.ExpertsFormat.MODULE_LIST:
        experts: hf.Rule = hf.ExpertsPerModule(
            hf_pattern="mlp.experts.{e}.{proj}.weight", native_pattern="mlp.grouped_experts.{proj}.weight",
            projections=(("gate_proj", "gate_proj"), ("up_proj", "up_proj"), ("down_proj", "down_proj")), num_experts=layer.num_experts)
    elif experts_format == hf.ExpertsFormat.FUSED:
        experts = hf.ExpertsFused(hf_gate_up="mlp.experts.gate_up_proj", hf_down="mlp.experts.down_proj",
                                  native_gate="mlp.grouped_experts.gate_proj.weight", native_up="mlp.grouped_experts.up_proj.weight",
                                  native_down="mlp.grouped_experts.down_proj.weight")
    else:
        raise ValueError(f"Unsupported experts format {experts_format}")
    shared = hf.shared_expert_rules() if layer.num_shared_experts > 0 else ()
    bias = (hf.Renamed("mlp.gate.e_score_correction_bias", "mlp.router.expert_bias"),) if layer.router.enable_expert_bias else ()
    return (hf.Renamed("mlp.gate.weight", "mlp.router.gate.weight"), *bias, experts, *shared)


def _backbone(params: DeepseekV2Parameters, experts_format: hf.ExpertsFormat) -> tuple[hf.Rule, ...]:
    def layer_rules(index: int) -> tuple[hf.Rule, ...]:
        return (*hf.latent_attention_rules(low_rank_query=params.layer.q_lora_rank is not None), *hf.norm_rules(),
                *_feed_forward(params.layer, index, experts_format))

    return hf.backbone_rules(layer_rules, params.num_hidden_layers, hf.single_vocab_name(params.split_vocab_order))


def mapper_from_huggingface_deepseek_v2(params: DeepseekV2Parameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(_backbone(params, experts_format), hf.Direction.FROM_HF)


def mapper_to_huggingface_deepseek_v2(params: DeepseekV2Parameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(_backbone(params, experts_format), hf.Direction.TO_HF)


def _causal(params: DeepseekV2ForCausalLMParameters, experts_format: hf.ExpertsFormat) -> tuple[hf.Rule, ...]:
    return hf.causal_lm_rules(_backbone(params.model, experts_format), hf.single_vocab_name(params.model.split_vocab_order))


def mapper_from_huggingface_deepseek_v2_for_causal_lm(params: DeepseekV2ForCausalLMParameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(_causal(params, experts_format), hf.Direction.FROM_HF)


def mapper_to_huggingface_deepseek_v2_for_causal_lm(params: DeepseekV2ForCausalLMParameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(_causal(params, experts_format), hf.Direction.TO_HF)


def mapper_from_huggingface_deepseek_v2_for_classification(params: DeepseekV2ForClassificationParameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(hf.classification_rules(_backbone(params.model, experts_format)), hf.Direction.FROM_HF)


def mapper_to_huggingface_deepseek_v2_for_classification(params: DeepseekV2ForClassificationParameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(hf.classification_rules(_backbone(params.model, experts_format)), hf.Direction.TO_HF)


def mapper_from_huggingface_deepseek_v2_for_embedding(params: DeepseekV2ForEmbeddingParameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(hf.embedding_rules(_backbone(params.model, experts_format)), hf.Direction.FROM_HF)


def mapper_to_huggingface_deepseek_v2_for_embedding(params: DeepseekV2ForEmbeddingParameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(hf.embedding_rules(_backbone(params.model, experts_format)), hf.Direction.TO_HF)
