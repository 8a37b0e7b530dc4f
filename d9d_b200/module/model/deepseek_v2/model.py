"""DeepSeek-V2 (net-new family: the reference ships the MLA / shared-expert blocks but no model using them)."""

from __future__ import annotations

from d9d_b200.module.block.hidden_states_aggregator import HiddenStatesAggregationMode
from d9d_b200.module.block.positional import RotaryEmbeddingStyle
from d9d_b200.module.model.decoder import DecoderBackbone, DecoderForCausalLM, DecoderForClassification, DecoderForEmbedding
from d9d_b200.pipelining.api import PipelineStageInfo

from .decoder_layer import DeepseekV2Layer
from .params import (
    DeepseekV2ForCausalLMParameters,
    DeepseekV2ForClassificationParameters,
    DeepseekV2ForEmbeddingParameters,
    DeepseekV2Parameters,
)


class DeepseekV2Model(DecoderBackbone):
    """Decoder backbone of the DeepSeek-V2 family, splittable across pipeline stages."""

    def __init__(self, params: DeepseekV2Parameters, stage: PipelineStageInfo,
                 hidden_states_snapshot_mode: HiddenStatesAggregationMode, enable_checkpointing: bool):
        super().__init__(params, stage, hidden_states_snapshot_mode, enable_checkpointing, layer_factory=DeepseekV2Layer,
                         rope_style=RotaryEmbeddingStyle.INTERLEAVED)


class DeepseekV2ForCausalLM(DecoderForCausalLM):
    def __init__(self, params: DeepseekV2ForCausalLMParameters, stage: PipelineStageInfo,
                 hidden_states_snapshot_mode: HiddenStatesAggregationMode, enable_checkpointing: bool):
        backbone = DeepseekV2Model(params.model, stage, hidden_states_snapshot_mode, enable_checkpointing)
        super().__init__(backbone, params.model, stage)


class DeepseekV2ForClassification(DecoderForClassification):
    def __init__(self, params: DeepseekV2ForClassificationParameters, stage: PipelineStageInfo,
                 hidden_states_snapshot_mode: HiddenStatesAggregationMode, enable_checkpointing: bool):
        backbone = DeepseekV2Model(params.model, stage, hidden_states_snapshot_mode, enable_checkpointing)
        super().__init__(backbone, params.model, stage, params.num_labels, params.classifier_dropout)


class DeepseekV2ForEmbedding(DecoderForEmbedding):
    def __init__(self, params: DeepseekV2ForEmbeddingParameters, stage: PipelineStageInfo,
                 hidden_states_snapshot_mode: HiddenStatesAggregationMode, enable_checkpointing: bool):
        backbone = DeepseekV2Model(params.model, stage, hidden_states_snapshot_mode, enable_checkpointing)
        super().__init__(backbone, params.model, stage, params.embedding_dim, params.normalize)
