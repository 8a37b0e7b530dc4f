"""Parallelisation plan of the Llama3 family (HSDP on dense units, EP on MoE layers)."""

from d9d_b200.core.dist_context import DistributedContext
from d9d_b200.module.model.llama3 import Llama3ForCausalLM, Llama3ForClassification, Llama3ForEmbedding, Llama3Model
from d9d_b200.pipelining.api import PipelineStageInfo

from ._plan import parallelize_backbone, parallelize_headed


def parallelize_llama3_model(dist_context: DistributedContext, model: Llama3Model, stage: PipelineStageInfo) -> None:
    parallelize_backbone(dist_context, model, stage)


def parallelize_llama3_for_causal_lm(dist_context: DistributedContext, model: Llama3ForCausalLM, stage: PipelineStageInfo) -> None:
    parallelize_headed(dist_context, model, stage, "lm_head")


def parallelize_llama3_for_classification(dist_context: DistributedContext, model: Llama3ForClassification, stage: PipelineStageInfo) -> None:
    parallelize_headed(dist_context, model, stage, "cls_head")


def parallelize_llama3_for_embedding(dist_context: DistributedContext, model: Llama3ForEmbedding, stage: PipelineStageInfo) -> None:
    parallelize_headed(dist_context, model, stage, "embedding_head")
