"""Parallelisation plan of the DeepSeek-V2 family (HSDP on dense units incl. latent attention, EP on MoE layers)."""

from d9d_b200.core.dist_context import DistributedContext
from d9d_b200.module.model.deepseek_v2 import DeepseekV2ForCausalLM, DeepseekV2ForClassification, DeepseekV2ForEmbedding, DeepseekV2Model
from d9d_b200.pipelining.api import PipelineStageInfo

from ._plan import parallelize_backbone, parallelize_headed


def parallelize_deepseek_v2_model(dist_context: DistributedContext, model: DeepseekV2Model, stage: PipelineStageInfo) -> None:
    parallelize_backbone(dist_context, model, stage)


def parallelize_deepseek_v2_for_causal_lm(dist_context: DistributedContext, model: DeepseekV2ForCausalLM, stage: PipelineStageInfo) -> None:
    parallelize_headed(dist_context, model, stage, "lm_head")


def parallelize_deepseek_v2_for_classification(dist_context: DistributedContext, model: DeepseekV2ForClassification, stage: PipelineStageInfo) -> None:
    parallelize_headed(dist_context, model, stage, "cls_head")


def parallelize_deepseek_v2_for_embedding(dist_context: DistributedContext, model: DeepseekV2ForEmbedding, stage: PipelineStageInfo) -> None:
    parallelize_headed(dist_context, model, stage, "embedding_head")
