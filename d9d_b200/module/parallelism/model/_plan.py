"""Shared plan: HSDP for dense sub-modules on the dense mesh, EP for MoE layers on the expert mesh, context-parallel
attention over the ``cp`` ranks of the batch mesh, tensor + sequence parallel attention / dense MLPs over ``tp``.

Parity: reference ``d9d/module/parallelism/model/qwen3_moe.py:12-146`` / ``qwen3_dense.py:12-145`` — every
sub-module (embeddings, final norm, per layer: attention, both norms, dense MLP, heads) is its own unit.
"""

from __future__ import annotations

from torch import nn

from d9d_b200.core.dist_context import BATCH_DOMAIN, DENSE_DOMAIN, EXPERT_DOMAIN, DistributedContext
from d9d_b200.module.block.attention import GroupedQueryAttention
from d9d_b200.module.block.ffn import SwiGLU
from d9d_b200.module.block.moe import MoELayer
from d9d_b200.module.model.decoder import DecoderBackbone
from d9d_b200.module.parallelism.api import (
    parallelize_attention_tensor_parallel,
    parallelize_context_parallel,
    parallelize_expert_parallel,
    parallelize_hsdp,
    parallelize_replicate,
    parallelize_swiglu_tensor_parallel,
)
from d9d_b200.pipelining.api import PipelineStageInfo

_DENSE_DIMS = ("dp_replicate", "dp_cp_shard", "cp_replicate")
_TP_DIMS = ("dp_replicate", "cp_replicate", "tp")  # mesh of the tensor-parallel linears; FSDP adds dp_cp_shard on top
_DENSE_TP_DIMS = ("dp_replicate", "dp_cp_shard", "cp_replicate", "tp")


def _check_supported(dist_context: DistributedContext) -> None:
    dims = dist_context.mesh_params
    sharded = dims.has_data_parallel_shard or dims.has_context_parallel_shard
    replicated = dims.has_data_parallel_replicate or dims.has_context_parallel_replicate
    if dims.has_tensor_parallel and sharded and replicated:
        # FSDP2 accepts tensor-parallel parameters only on a 1-D ``tp`` mesh, so they cannot also carry replicate dims
        raise ValueError("Tensor parallelism combines with FSDP sharding (data_parallel_shard / context_parallel_shard) or with "
                         "replication (data_parallel_replicate / context_parallel_replicate), not with both at once.")


def _tensor_parallel_mesh(dist_context: DistributedContext):  # noqa: ANN202
    dense = dist_context.mesh_for(DENSE_DOMAIN)
    return dense[tuple(d for d in _TP_DIMS if d == "tp" or dense[d].size() > 1)]


def dense_unit(dist_context: DistributedContext, module: nn.Module) -> None:
    """HSDP over the data / context dims; with tensor parallelism the unit is additionally replicated over ``tp`` (its
    activations are sequence-sharded there, so its gradients are partial sums that must be reduced over ``tp`` too)."""
    dense = dist_context.mesh_for(DENSE_DOMAIN)
    if dist_context.mesh_params.has_tensor_parallel:
        parallelize_hsdp(module, mesh=dense[_DENSE_TP_DIMS], skip_distributed=True)
    else:
        parallelize_hsdp(module, mesh=dense[_DENSE_DIMS])


def _tensor_parallel_layer(dist_context: DistributedContext, layer: nn.Module) -> None:
    """Megatron-style tensor + sequence parallelism of one decoder layer: hidden states between blocks carry ``S / tp``
    tokens; attention and the dense MLP gather them in their column-parallel input GEMMs and scatter them again in their
    row-parallel output GEMMs (fused into the GEMMs on CUDA)."""
    mesh = _tensor_parallel_mesh(dist_context)
    if not isinstance(layer.self_attn, GroupedQueryAttention):
        raise ValueError(f"tensor parallelism is implemented for GroupedQueryAttention, got {type(layer.self_attn).__name__}")
    parallelize_attention_tensor_parallel(layer.self_attn, mesh)
    if isinstance(layer.mlp, SwiGLU):
        parallelize_swiglu_tensor_parallel(layer.mlp, mesh)


def parallelize_backbone(dist_context: DistributedContext, model: DecoderBackbone, stage: PipelineStageInfo) -> None:
    _check_supported(dist_context)
    expert_mesh = dist_context.mesh_for(EXPERT_DOMAIN)["ep_replicate", "ep_shard"]
    # cp_shard and cp_replicate both split the sequence (they differ in how the weights are held): attention spans both
    cp_mesh = dist_context.mesh_for(BATCH_DOMAIN)["cp"]
    tensor_parallel = dist_context.mesh_params.has_tensor_parallel
    if stage.is_current_stage_first:
        dense_unit(dist_context, model.embed_tokens)
    if stage.is_current_stage_last:
        dense_unit(dist_context, model.norm)
    for layer in model.layers.values():
        if tensor_parallel:
            _tensor_parallel_layer(dist_context, layer)
        if isinstance(layer.mlp, MoELayer):
            parallelize_expert_parallel(layer.mlp, mesh_experts=expert_mesh)
        else:
            dense_unit(dist_context, layer.mlp)  # with tensor parallelism: FSDP-shards the column / row parallel weights
        if cp_mesh.size() > 1:
            parallelize_context_parallel(layer.self_attn, cp_mesh)
        dense_unit(dist_context, layer.self_attn)  # with tensor parallelism: what the TP styles left (q/k norms)
        dense_unit(dist_context, layer.input_layernorm)
        dense_unit(dist_context, layer.post_attention_layernorm)


def parallelize_headed(dist_context: DistributedContext, model: nn.Module, stage: PipelineStageInfo, head_attr: str) -> None:
    parallelize_backbone(dist_context, model.model, stage)
    if stage.is_current_stage_last:
        dense_unit(dist_context, getattr(model, head_attr))
