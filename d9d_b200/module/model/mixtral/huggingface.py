"""HuggingFace ⇄ native state mappers of the Mixtral family (rules in ``module/model/_huggingface.py``)."""

from __future__ import annotations

from d9d_b200.model_state.mapper import ModelStateMapper
from d9d_b200.module.model import _huggingface as hf

from .params import (
    MixtralForCausalLMParameters,
    MixtralForClassificationParameters,
    MixtralForEmbeddingParameters,
    MixtralLayerParameters,
    MixtralParameters,
)

MixtralExpertsFormat = hf.ExpertsFormat


def _experts(layer: MixtralLayerParameters, experts_format: hf.ExpertsFormat) -> tuple[hf.Rule, ...]:
    if experts_format == hf.ExpertsFormat.MODULE_LIST:  # Mixtral names: w1 = gate, w3 = up, w2 = down
        return (hf.ExpertsPerModule(hf_pattern="block_sparse_moe.experts.{e}.{proj}.weight",
                                    projections=(("w1", "gate_proj"), ("w3", "up_proj"), ("w2", "down_proj")),
                                    native_pattern="mlp.grouped_experts.{proj}.weight", num_experts=layer.num_experts),)
    if experts_format == hf.ExpertsFormat.FUSED:
        return (hf.ExpertsFused(hf_gate_up="mlp.experts.gate_up_proj", hf_down="mlp.experts.down_proj",
                                native_gate="mlp.grouped_experts.gate_proj.weight", native_up="mlp.grouped_experts.up_proj.weight",
                                native_down="mlp.grouped_experts.down_proj.weight"),)
    raise ValueError(f"Unsupported experts format {experts_format}")


def _backbone(params: MixtralParameters, experts_format: hf.ExpertsFormat) -> tuple[hf.Rule, ...]:
    router = (hf.Renamed("block_sparse_moe.gate.weight", "mlp.router.gate.weight") if experts_format == hf.ExpertsFormat.MODULE_LIST
              else hf.Renamed("mlp.gate.weight", "mlp.router.gate.weight"))
    layer = (*hf.attention_rules(qk_norm=False), *hf.norm_rules(), router, *_experts(params.layer, experts_format))
    return hf.backbone_rules(layer, params.num_hidden_layers, hf.single_vocab_name(params.split_vocab_order))


def mapper_from_huggingface_mixtral(params: MixtralParameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(_backbone(params, experts_format), hf.Direction.FROM_HF)


def mapper_to_huggingface_mixtral(params: MixtralParameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(_backbone(params, experts_format), hf.Direction.TO_HF)


def _causal(params: MixtralForCausalLMParameters, experts_format: hf.ExpertsFormat) -> tuple[hf.Rule, ...]:
    return hf.causal_lm_rules(_backbone(params.model, experts_format), hf.single_vocab_name(params.model.split_vocab_order))


def mapper_from_huggingface_mixtral_for_causal_lm(params: MixtralForCausalLMParameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(_causal(params, experts_format), hf.Direction.FROM_HF)


def mapper_to_huggingface_mixtral_for_causal_lm(params: MixtralForCausalLMParameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(_causal(params, experts_format), hf.Direction.TO_HF)


def mapper_from_huggingface_mixtral_for_classification(params: MixtralForClassificationParameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(hf.classification_rules(_backbone(params.model, experts_format)), hf.Direction.FROM_HF)


def mapper_to_huggingface_mixtral_for_classification(params: MixtralForClassificationParameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(hf.classification_rules(_backbone(params.model, experts_format)), hf.Direction.TO_HF)


def mapper_from_huggingface_mixtral_for_embedding(params: MixtralForEmbeddingParameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(hf.embedding_rules(_backbone(params.model, experts_format)), hf.Direction.FROM_HF)


def mapper_to_huggingface_mixtral_for_embedding(params: MixtralForEmbeddingParameters, experts_format: hf.ExpertsFormat) -> ModelStateMapper:
    return hf.compile_rules(hf.embedding_rules(_backbone(params.model, experts_format)), hf.Direction.TO_HF)
