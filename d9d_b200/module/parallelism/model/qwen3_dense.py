"""Parallelisation plan of the Qwen3Dense family (HSDP on dense units, EP on MoE layers)."""

from d9d_b200.core.dist_context import DistributedContext
from d9d_b200.module.model.qwen3_dense import Qwen3DenseForCausalLM, Qwen3DenseForClassification, Qwen3DenseForEmbedding, Qwen3DenseModel
from d9d_b200.pipelining.api import PipelineStageInfo

from ._plan import parallelize_backbone, parallelize_headed


def parallelize_qwen3_dense_model(dist_context: DistributedContext, model: Qwen3DenseModel, stage: PipelineStageInfo) -> None:
    parallelize_backbone(dist_context, model, stage)


def parallelize_qwen3_dense_for_causal_lm(dist_context: DistributedContext, model: Qwen3DenseForCausalLM, stage: PipelineStageInfo) -> None:
    parallelize_headed(dist_context, model, stage, "lm_head")


def parallelize_qwen3_dense_for_classification(dist_context: DistributedContext, model: Qwen3DenseForClassification, stage: PipelineStageInfo) -> None:
    parallelize_headed(dist_context, model, stage, "cls_head")


def parallelize_qwen3_dense_for_embedding(dist_context: DistributedContext, model: Qwen3DenseForEmbedding, stage: PipelineStageInfo) -> None:
    parallelize_headed(dist_context, model, stage, "embedding_head")
