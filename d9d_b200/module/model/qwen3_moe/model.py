"""Qwen3 Mixture-of-Experts (reference ``d9d/module/model/qwen3_moe``)."""

from __future__ import annotations

from d9d_b200.module.block.hidden_states_aggregator import HiddenStatesAggregationMode
from d9d_b200.module.model.decoder import DecoderBackbone, DecoderForCausalLM, DecoderForClassification, DecoderForEmbedding
from d9d_b200.pipelining.api import PipelineStageInfo

from .decoder_layer import Qwen3MoELayer
from .params import (
    Qwen3MoEForCausalLMParameters,
    Qwen3MoEForClassificationParameters,
    Qwen3MoEForEmbeddingParameters,
    Qwen3MoEParameters,
)


class Qwen3MoEModel(DecoderBackbone):
    """Decoder backbone of the Qwen3MoE family, splittable across pipeline stages."""

    def __init__(self, params: Qwen3MoEParameters, stage: PipelineStageInfo,
                 hidden_states_snapshot_mode: HiddenStatesAggregationMode, enable_checkpointing: bool):
        super().__init__(params, stage, hidden_states_snapshot_mode, enable_checkpointing, layer_factory=Qwen3MoELayer)


class Qwen3MoEForCausalLM(DecoderForCausalLM):
    def __init__(self, params: Qwen3MoEForCausalLMParameters, stage: PipelineStageInfo,
                 hidden_states_snapshot_mode: HiddenStatesAggregationMode, enable_checkpointing: bool):
        backbone = Qwen3MoEModel(params.model, stage, hidden_states_snapshot_mode, enable_checkpointing)
        super().__init__(backbone, params.model, stage)


class Qwen3MoEForClassification(DecoderForClassification):
    def __init__(self, params: Qwen3MoEForClassificationParameters, stage: PipelineStageInfo,
                 hidden_states_snapshot_mode: HiddenStatesAggregationMode, enable_checkpointing: bool):
        backbone = Qwen3MoEModel(params.model, stage, hidden_states_snapshot_mode, enable_checkpointing)
        super().__init__(backbone, params.model, stage, params.num_labels, params.classifier_dropout)


class Qwen3MoEForEmbedding(DecoderForEmbedding):
    def __init__(self, params: Qwen3MoEForEmbeddingParameters, stage: PipelineStageInfo,
                 hidden_states_snapshot_mode: HiddenStatesAggregationMode, enable_checkpointing: bool):
        backbone = Qwen3MoEModel(params.model, stage, hidden_states_snapshot_mode, enable_checkpointing)
        super().__init__(backbone, params.model, stage, params.embedding_dim, params.normalize)
