"""Mixtral style sparse-MoE decoder (no q/k norm) — net-new family built from the shared blocks."""

from __future__ import annotations

from d9d_b200.module.block.hidden_states_aggregator import HiddenStatesAggregationMode
from d9d_b200.module.model.decoder import DecoderBackbone, DecoderForCausalLM, DecoderForClassification, DecoderForEmbedding
from d9d_b200.pipelining.api import PipelineStageInfo

from .decoder_layer import MixtralLayer
from .params import (
    MixtralForCausalLMParameters,
    MixtralForClassificationParameters,
    MixtralForEmbeddingParameters,
    MixtralParameters,
)


class MixtralModel(DecoderBackbone):
    """Decoder backbone of the Mixtral family, splittable across pipeline stages."""

    def __init__(self, params: MixtralParameters, stage: PipelineStageInfo,
                 hidden_states_snapshot_mode: HiddenStatesAggregationMode, enable_checkpointing: bool):
        super().__init__(params, stage, hidden_states_snapshot_mode, enable_checkpointing, layer_factory=MixtralLayer)


class MixtralForCausalLM(DecoderForCausalLM):
    def __init__(self, params: MixtralForCausalLMParameters, stage: PipelineStageInfo,
                 hidden_states_snapshot_mode: HiddenStatesAggregationMode, enable_checkpointing: bool):
        backbone = MixtralModel(params.model, stage, hidden_states_snapshot_mode, enable_checkpointing)
        super().__init__(backbone, params.model, stage)


class MixtralForClassification(DecoderForClassification):
    def __init__(self, params: MixtralForClassificationParameters, stage: PipelineStageInfo,
                 hidden_states_snapshot_mode: HiddenStatesAggregationMode, enable_checkpointing: bool):
        backbone = MixtralModel(params.model, stage, hidden_states_snapshot_mode, enable_checkpointing)
        super().__init__(backbone, params.model, stage, params.num_labels, params.classifier_dropout)


class MixtralForEmbedding(DecoderForEmbedding):
    def __init__(self, params: MixtralForEmbeddingParameters, stage: PipelineStageInfo,
                 hidden_states_snapshot_mode: HiddenStatesAggregationMode, enable_checkpointing: bool):
        backbone = MixtralModel(params.model, stage, hidden_states_snapshot_mode, enable_checkpointing)
        super().__init__(backbone, params.model, stage, params.embedding_dim, params.normalize)
