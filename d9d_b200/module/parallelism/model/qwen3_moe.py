"""Parallelisation plan of the Qwen3MoE family (HSDP on dense units, EP on MoE layers)."""

from d9d_b200.core.dist_context import DistributedContext
from d9d_b200.module.model.qwen3_moe import Qwen3MoEForCausalLM, Qwen3MoEForClassification, Qwen3MoEForEmbedding, Qwen3MoEModel
from d9d_b200.pipelining.api import PipelineStageInfo

from ._plan import parallelize_backbone, parallelize_headed


def parallelize_qwen3_moe_model(dist_context: DistributedContext, model: Qwen3MoEModel, stage: PipelineStageInfo) -> None:
    parallelize_backbone(dist_context, model, stage)


def parallelize_qwen3_moe_for_causal_lm(dist_context: DistributedContext, model: Qwen3MoEForCausalLM, stage: PipelineStageInfo) -> None:
    parallelize_headed(dist_context, model, stage, "lm_head")


def parallelize_qwen3_moe_for_classification(dist_context: DistributedContext, model: Qwen3MoEForClassification, stage: PipelineStageInfo) -> None:
    parallelize_headed(dist_context, model, stage, "cls_head")


def parallelize_qwen3_moe_for_embedding(dist_context: DistributedContext, model: Qwen3MoEForEmbedding, stage: PipelineStageInfo) -> None:
    parallelize_headed(dist_context, model, stage, "embedding_head")
