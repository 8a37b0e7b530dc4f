"""Shared plan: HSDP for dense sub-modules on the dense mesh, EP for MoE layers on the expert mesh, context-parallel
attention over the ``cp`` ranks of the batch mesh.

Parity: reference ``d9d/module/parallelism/model/qwen3_moe.py:12-146`` / ``qwen3_dense.py:12-145`` — every
sub-module (embeddings, final norm, per layer: attention, both norms, dense MLP, heads) is its own unit.
"""

from __future__ import annotations

from torch import nn

from d9d_b200.core.dist_context import BATCH_DOMAIN, DENSE_DOMAIN, EXPERT_DOMAIN, DistributedContext
from d9d_b200.module.block.moe import MoELayer
from d9d_b200.module.model.decoder import DecoderBackbone
from d9d_b200.module.parallelism.api import parallelize_context_parallel, parallelize_expert_parallel, parallelize_hsdp
from d9d_b200.pipelining.api import PipelineStageInfo

_DENSE_DIMS = ("dp_replicate", "dp_cp_shard", "cp_replicate")


def _check_supported(dist_context: DistributedContext) -> None:
    dims = dist_context.mesh_params
    if dims.has_tensor_parallel:
        raise ValueError("Tensor Parallel currently is not supported for this model.")


def dense_unit(dist_context: DistributedContext, module: nn.Module) -> None:
    parallelize_hsdp(module, mesh=dist_context.mesh_for(DENSE_DOMAIN)[_DENSE_DIMS])


def parallelize_backbone(dist_context: DistributedContext, model: DecoderBackbone, stage: PipelineStageInfo) -> None:
    _check_supported(dist_context)
    expert_mesh = dist_context.mesh_for(EXPERT_DOMAIN)["ep_replicate", "ep_shard"]
    # cp_shard and cp_replicate both split the sequence (they differ in how the weights are held): attention spans both
    cp_mesh = dist_context.mesh_for(BATCH_DOMAIN)["cp"]
    if stage.is_current_stage_first:
        dense_unit(dist_context, model.embed_tokens)
    if stage.is_current_stage_last:
        dense_unit(dist_context, model.norm)
    for layer in model.layers.values():
        if isinstance(layer.mlp, MoELayer):
            parallelize_expert_parallel(layer.mlp, mesh_experts=expert_mesh)
        else:
            dense_unit(dist_context, layer.mlp)
        if cp_mesh.size() > 1:
            parallelize_context_parallel(layer.self_attn, cp_mesh)
        dense_unit(dist_context, layer.self_attn)
        dense_unit(dist_context, layer.input_layernorm)
        dense_unit(dist_context, layer.post_attention_layernorm)


def parallelize_headed(dist_context: DistributedContext, model: nn.Module, stage: PipelineStageInfo, head_attr: str) -> None:
    parallelize_backbone(dist_context, model.model, stage)
    if stage.is_current_stage_last:
        dense_unit(dist_context, getattr(model, head_attr))
