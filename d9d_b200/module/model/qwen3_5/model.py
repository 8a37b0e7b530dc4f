"""Qwen3.5 dense text model (net-new family: the reference ships the Gated DeltaNet / gated attention blocks but no
model using them)."""

from __future__ import annotations

from d9d_b200.module.block.hidden_states_aggregator import HiddenStatesAggregationMode
from d9d_b200.module.model.decoder import DecoderBackbone, DecoderForCausalLM, DecoderForClassification, DecoderForEmbedding
from d9d_b200.pipelining.api import PipelineStageInfo

from .decoder_layer import Qwen3_5Layer
from .params import (
    Qwen3_5ForCausalLMParameters,
    Qwen3_5ForClassificationParameters,
    Qwen3_5ForEmbeddingParameters,
    Qwen3_5Parameters,
)


class Qwen3_5Model(DecoderBackbone):
    """Hybrid linear / softmax attention backbone, splittable across pipeline stages."""

    def __init__(self, params: Qwen3_5Parameters, stage: PipelineStageInfo,
                 hidden_states_snapshot_mode: HiddenStatesAggregationMode, enable_checkpointing: bool):
        super().__init__(params, stage, hidden_states_snapshot_mode, enable_checkpointing, layer_factory=Qwen3_5Layer,
                         zero_centered_norm=True)


class Qwen3_5ForCausalLM(DecoderForCausalLM):
    def __init__(self, params: Qwen3_5ForCausalLMParameters, stage: PipelineStageInfo,
                 hidden_states_snapshot_mode: HiddenStatesAggregationMode, enable_checkpointing: bool):
        backbone = Qwen3_5Model(params.model, stage, hidden_states_snapshot_mode, enable_checkpointing)
        super().__init__(backbone, params.model, stage)


class Qwen3_5ForClassification(DecoderForClassification):
    def __init__(self, params: Qwen3_5ForClassificationParameters, stage: PipelineStageInfo,
                 hidden_states_snapshot_mode: HiddenStatesAggregationMode, enable_checkpointing: bool):
        backbone = Qwen3_5Model(params.model, stage, hidden_states_snapshot_mode, enable_checkpointing)
        super().__init__(backbone, params.model, stage, params.num_labels, params.classifier_dropout)


class Qwen3_5ForEmbedding(DecoderForEmbedding):
    def __init__(self, params: Qwen3_5ForEmbeddingParameters, stage: PipelineStageInfo,
                 hidden_states_snapshot_mode: HiddenStatesAggregationMode, enable_checkpointing: bool):
        backbone = Qwen3_5Model(params.model, stage, hidden_states_snapshot_mode, enable_checkpointing)
        super().__init__(backbone, params.model, stage, params.embedding_dim, params.normalize)
