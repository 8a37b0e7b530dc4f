"""Qwen3 dense (reference ``d9d/module/model/qwen3_dense``)."""

from __future__ import annotations

from d9d_b200.module.block.hidden_states_aggregator import HiddenStatesAggregationMode
from d9d_b200.module.model.decoder import DecoderBackbone, DecoderForCausalLM, DecoderForClassification, DecoderForEmbedding
from d9d_b200.pipelining.api import PipelineStageInfo

from .decoder_layer import Qwen3DenseLayer
from .params import (
    Qwen3DenseForCausalLMParameters,
    Qwen3DenseForClassificationParameters,
    Qwen3DenseForEmbeddingParameters,
    Qwen3DenseParameters,
)


class Qwen3DenseModel(DecoderBackbone):
    """Decoder backbone of the Qwen3Dense family, splittable across pipeline stages."""

    def __init__(self, params: Qwen3DenseParameters, stage: PipelineStageInfo,
                 hidden_states_snapshot_mode: HiddenStatesAggregationMode, enable_checkpointing: bool):
        super().__init__(params, stage, hidden_states_snapshot_mode, enable_checkpointing, layer_factory=Qwen3DenseLayer)


class Qwen3DenseForCausalLM(DecoderForCausalLM):
    def __init__(self, params: Qwen3DenseForCausalLMParameters, stage: PipelineStageInfo,
                 hidden_states_snapshot_mode: HiddenStatesAggregationMode, enable_checkpointing: bool):
        backbone = Qwen3DenseModel(params.model, stage, hidden_states_snapshot_mode, enable_checkpointing)
        super().__init__(backbone, params.model, stage)


class Qwen3DenseForClassification(DecoderForClassification):
    def __init__(self, params: Qwen3DenseForClassificationParameters, stage: PipelineStageInfo,
                 hidden_states_snapshot_mode: HiddenStatesAggregationMode, enable_checkpointing: bool):
        backbone = Qwen3DenseModel(params.model, stage, hidden_states_snapshot_mode, enable_checkpointing)
        super().__init__(backbone, params.model, stage, params.num_labels, params.classifier_dropout)


class Qwen3DenseForEmbedding(DecoderForEmbedding):
    def __init__(self, params: Qwen3DenseForEmbeddingParameters, stage: PipelineStageInfo,
                 hidden_states_snapshot_mode: HiddenStatesAggregationMode, enable_checkpointing: bool):
        backbone = Qwen3DenseModel(params.model, stage, hidden_states_snapshot_mode, enable_checkpointing)
        super().__init__(backbone, params.model, stage, params.embedding_dim, params.normalize)
