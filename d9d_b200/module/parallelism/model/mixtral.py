"""Parallelisation plan of the Mixtral family (HSDP on dense units, EP on MoE layers)."""

from d9d_b200.core.dist_context import DistributedContext
from d9d_b200.module.model.mixtral import MixtralForCausalLM, MixtralForClassification, MixtralForEmbedding, MixtralModel
from d9d_b200.pipelining.api import PipelineStageInfo

from ._plan import parallelize_backbone, parallelize_headed


def parallelize_mixtral_model(dist_context: DistributedContext, model: MixtralModel, stage: PipelineStageInfo) -> None:
    parallelize_backbone(dist_context, model, stage)


def parallelize_mixtral_for_causal_lm(dist_context: DistributedContext, model: MixtralForCausalLM, stage: PipelineStageInfo) -> None:
    parallelize_headed(dist_context, model, stage, "lm_head")


def parallelize_mixtral_for_classification(dist_context: DistributedContext, model: MixtralForClassification, stage: PipelineStageInfo) -> None:
    parallelize_headed(dist_context, model, stage, "cls_head")


def parallelize_mixtral_for_embedding(dist_context: DistributedContext, model: MixtralForEmbedding, stage: PipelineStageInfo) -> None:
    parallelize_headed(dist_context, model, stage, "embedding_head")
