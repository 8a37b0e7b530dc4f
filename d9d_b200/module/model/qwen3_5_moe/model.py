"""Qwen3.5-MoE text model (net-new family)."""

from __future__ import annotations

from d9d_b200.module.block.hidden_states_aggregator import HiddenStatesAggregationMode
from d9d_b200.module.block.moe import MoELayer, SharedExpertParameters
from d9d_b200.module.model.decoder import (
    DecoderBackbone,
    DecoderForCausalLM,
    DecoderForClassification,
    DecoderForEmbedding,
    PreNormDecoderLayer,
)
from d9d_b200.module.model.qwen3_5.decoder_layer import build_token_mixer
from d9d_b200.pipelining.api import PipelineStageInfo

from .params import (
    Qwen3_5MoEForCausalLMParameters,
    Qwen3_5MoEForClassificationParameters,
    Qwen3_5MoEForEmbeddingParameters,
    Qwen3_5MoELayerParameters,
    Qwen3_5MoEParameters,
)


class Qwen3_5MoELayer(PreNormDecoderLayer):
    """Hybrid token mixer (see ``qwen3_5``) + MoE feed-forward with a sigmoid-gated shared expert; zero-centred RMSNorms."""

    def __init__(self, params: Qwen3_5MoELayerParameters, index: int):
        mlp = MoELayer(hidden_dim=params.hidden_size, intermediate_dim_grouped=params.moe_intermediate_size,
                       num_grouped_experts=params.num_experts, top_k=params.experts_top_k, router_renormalize_probabilities=True,
                       shared_expert=SharedExpertParameters(intermediate_size=params.shared_expert_intermediate_size, enable_gate=True))
        super().__init__(build_token_mixer(params, index), mlp, params.hidden_size, params.rms_norm_eps, zero_centered_norm=True)


class Qwen3_5MoEModel(DecoderBackbone):
    def __init__(self, params: Qwen3_5MoEParameters, stage: PipelineStageInfo,
                 hidden_states_snapshot_mode: HiddenStatesAggregationMode, enable_checkpointing: bool):
        super().__init__(params, stage, hidden_states_snapshot_mode, enable_checkpointing, layer_factory=Qwen3_5MoELayer,
                         zero_centered_norm=True)


class Qwen3_5MoEForCausalLM(DecoderForCausalLM):
    def __init__(self, params: Qwen3_5MoEForCausalLMParameters, stage: PipelineStageInfo,
                 hidden_states_snapshot_mode: HiddenStatesAggregationMode, enable_checkpointing: bool):
        backbone = Qwen3_5MoEModel(params.model, stage, hidden_states_snapshot_mode, enable_checkpointing)
        super().__init__(backbone, params.model, stage)


class Qwen3_5MoEForClassification(DecoderForClassification):
    def __init__(self, params: Qwen3_5MoEForClassificationParameters, stage: PipelineStageInfo,
                 hidden_states_snapshot_mode: HiddenStatesAggregationMode, enable_checkpointing: bool):
        backbone = Qwen3_5MoEModel(params.model, stage, hidden_states_snapshot_mode, enable_checkpointing)
        super().__init__(backbone, params.model, stage, params.num_labels, params.classifier_dropout)


class Qwen3_5MoEForEmbedding(DecoderForEmbedding):
    def __init__(self, params: Qwen3_5MoEForEmbeddingParameters, stage: PipelineStageInfo,
                 hidden_states_snapshot_mode: HiddenStatesAggregationMode, enable_checkpointing: bool):
        backbone = Qwen3_5MoEModel(params.model, stage, hidden_states_snapshot_mode, enable_checkpointing)
        super().__init__(backbone, params.model, stage, params.embedding_dim, params.normalize)
