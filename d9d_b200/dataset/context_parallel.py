from __future__ import annotations

from typing import Any, TypeVar

import torch
from torch.utils import _pytree as pytree

from d9d_b200.core.dist_context import BATCH_DOMAIN, DistributedContext
from d9d_b200.kernel.context_parallel import ContextParallelLayout, shard_sequence

T = TypeVar("T")


def context_parallel_rank_and_size(dist_context: DistributedContext) -> tuple[int, int]:
    """This rank's index in / the size of its context-parallel group (``(0, 1)`` without context parallelism)."""
    if not dist_context.mesh_params.is_distributed:
        return 0, 1
    cp = dist_context.mesh_for(BATCH_DOMAIN)["cp"]
    return cp.get_local_rank(), cp.size()


def shard_batch_for_context_parallel(batch: T, dist_context: DistributedContext, seq_dim: int = 1,
                                     layout: ContextParallelLayout = ContextParallelLayout.zigzag) -> T:
    """Keep only this context-parallel rank's tokens of every tensor in ``batch`` (any pytree; tensors with fewer than
    ``seq_dim + 1`` dims and non-tensors are returned unchanged).

    Call it in ``Task.build_forward_inputs`` on the collated batch - ``input_ids``, ``position_ids`` (which keep their
    *global* values, so rotary embeddings stay correct), ``labels`` and masks are cut consistently.  Ranks of one
    context-parallel group read the same samples (the data-parallel sharding of datasets ignores the ``cp`` dim), each
    keeps ``S / cp`` tokens of every sequence; losses weighted by the local token count then combine to the exact global
    mean.  Identity when the context-parallel degree is 1.
    """
    rank, world = context_parallel_rank_and_size(dist_context)
    return _cut(batch, seq_dim, world, rank, layout)


def _cut(batch: T, seq_dim: int, world: int, rank: int, layout: ContextParallelLayout) -> T:
    if world == 1:
        return batch

    def cut(leaf: Any) -> Any:
        if isinstance(leaf, torch.Tensor) and leaf.ndim > seq_dim:
            return shard_sequence(leaf, seq_dim, world, rank, layout)
        return leaf

    return pytree.tree_map(cut, batch)


def shard_batch_for_sequence_parallel(batch: T, dist_context: DistributedContext, seq_dim: int = 1) -> T:
    """Keep this *tensor-parallel* rank's contiguous chunk of every sequence: the family plans run tensor parallelism
    with sequence parallelism, i.e. activations between attention / MLP blocks carry ``S / tp`` tokens (the blocks gather
    and scatter them in their GEMMs).  Identity when the tensor-parallel degree is 1."""
    if not dist_context.mesh_params.is_distributed:
        return batch
    tp = dist_context.mesh_for(BATCH_DOMAIN)["tp"]
    return _cut(batch, seq_dim, tp.size(), tp.get_local_rank(), ContextParallelLayout.contiguous)


def shard_batch_along_sequence(batch: T, dist_context: DistributedContext, seq_dim: int = 1,
                               layout: ContextParallelLayout = ContextParallelLayout.zigzag) -> T:
    """Context-parallel cut (``layout``) followed by the tensor/sequence-parallel cut: what a task should apply to the
    collated batch so that any mesh works.  Identity for meshes without ``cp`` / ``tp``."""
    return shard_batch_for_sequence_parallel(shard_batch_for_context_parallel(batch, dist_context, seq_dim, layout),
                                             dist_context, seq_dim)
