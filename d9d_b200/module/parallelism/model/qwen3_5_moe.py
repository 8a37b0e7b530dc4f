"""Parallelisation plan of the Qwen3.5-MoE family (HSDP on dense units, expert parallel MoE layers; linear-attention layers cannot be context- or
tensor-parallel: their recurrence would need state passing between ranks)."""

from d9d_b200.core.dist_context import DistributedContext
from d9d_b200.module.model.qwen3_5_moe import Qwen3_5MoEForCausalLM, Qwen3_5MoEForClassification, Qwen3_5MoEForEmbedding, Qwen3_5MoEModel
from d9d_b200.pipelining.api import PipelineStageInfo

from ._plan import parallelize_backbone, parallelize_headed


def _check(dist_context: DistributedContext) -> None:
    dims = dist_context.mesh_params
    if dims.has_tensor_parallel or dims.has_context_parallel_shard or dims.has_context_parallel_replicate:
        raise ValueError("Gated DeltaNet layers do not support context / tensor parallelism yet.")


def parallelize_qwen3_5_moe_model(dist_context: DistributedContext, model: Qwen3_5MoEModel, stage: PipelineStageInfo) -> None:
    _check(dist_context)
    parallelize_backbone(dist_context, model, stage)


def parallelize_qwen3_5_moe_for_causal_lm(dist_context: DistributedContext, model: Qwen3_5MoEForCausalLM, stage: PipelineStageInfo) -> None:
    _check(dist_context)
    parallelize_headed(dist_context, model, stage, "lm_head")


def parallelize_qwen3_5_moe_for_classification(dist_context: DistributedContext, model: Qwen3_5MoEForClassification, stage: PipelineStageInfo) -> None:
    _check(dist_context)
    parallelize_headed(dist_context, model, stage, "cls_head")


def parallelize_qwen3_5_moe_for_embedding(dist_context: DistributedContext, model: Qwen3_5MoEForEmbedding, stage: PipelineStageInfo) -> None:
    _check(dist_context)
    parallelize_headed(dist_context, model, stage, "embedding_head")
