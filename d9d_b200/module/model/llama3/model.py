"""Llama-3 style dense decoder (no q/k norm) — net-new family built from the shared blocks."""

from __future__ import annotations

from d9d_b200.module.block.hidden_states_aggregator import HiddenStatesAggregationMode
from d9d_b200.module.model.decoder import DecoderBackbone, DecoderForCausalLM, DecoderForClassification, DecoderForEmbedding
from d9d_b200.pipelining.api import PipelineStageInfo

from .decoder_layer import Llama3Layer
from .params import (
    Llama3ForCausalLMParameters,
    Llama3ForClassificationParameters,
    Llama3ForEmbeddingParameters,
    Llama3Parameters,
)


class Llama3Model(DecoderBackbone):
    """Decoder backbone of the Llama3 family, splittable across pipeline stages."""

    def __init__(self, params: Llama3Parameters, stage: PipelineStageInfo,
                 hidden_states_snapshot_mode: HiddenStatesAggregationMode, enable_checkpointing: bool):
        super().__init__(params, stage, hidden_states_snapshot_mode, enable_checkpointing, layer_factory=Llama3Layer)


class Llama3ForCausalLM(DecoderForCausalLM):
    def __init__(self, params: Llama3ForCausalLMParameters, stage: PipelineStageInfo,
                 hidden_states_snapshot_mode: HiddenStatesAggregationMode, enable_checkpointing: bool):
        backbone = Llama3Model(params.model, stage, hidden_states_snapshot_mode, enable_checkpointing)
        super().__init__(backbone, params.model, stage)


class Llama3ForClassification(DecoderForClassification):
    def __init__(self, params: Llama3ForClassificationParameters, stage: PipelineStageInfo,
                 hidden_states_snapshot_mode: HiddenStatesAggregationMode, enable_checkpointing: bool):
        backbone = Llama3Model(params.model, stage, hidden_states_snapshot_mode, enable_checkpointing)
        super().__init__(backbone, params.model, stage, params.num_labels, params.classifier_dropout)


class Llama3ForEmbedding(DecoderForEmbedding):
    def __init__(self, params: Llama3ForEmbeddingParameters, stage: PipelineStageInfo,
                 hidden_states_snapshot_mode: HiddenStatesAggregationMode, enable_checkpointing: bool):
        backbone = Llama3Model(params.model, stage, hidden_states_snapshot_mode, enable_checkpointing)
        super().__init__(backbone, params.model, stage, params.embedding_dim, params.normalize)
