from __future__ import annotations

from collections.abc import Sequence
from typing import Any, TypeVar

import torch
import torch.utils._pytree as pytree

from .spec import ShardingSpec, ShardingSpecLeaf, SpecReplicate, SpecShard

TTree = TypeVar("TTree")


def _is_data_leaf(x: Any) -> bool:
    # lists are data (sequences of per-sample python objects), not containers
    return isinstance(x, (torch.Tensor, list))


# ------------------------------------------------------------------------------------------------ spec inference
def shard_spec_on_dim(tree: Any, dim: int) -> ShardingSpec:
    """Spec that splits every tensor on ``dim`` (lists on 0), replicating scalars and non-tensor leaves."""

    def infer(leaf: Any) -> ShardingSpecLeaf:
        if isinstance(leaf, list):
            if dim != 0:
                raise ValueError(f"lists are one-dimensional and can only be sharded on dim 0 (asked for {dim})")
            return SpecShard(0)
        if isinstance(leaf, torch.Tensor):
            if leaf.ndim == 0:
                return SpecReplicate()
            if dim >= leaf.ndim:
                raise ValueError(f"cannot shard a {leaf.ndim}-d tensor on dim {dim}")
            return SpecShard(dim)
        return SpecReplicate()

    return pytree.tree_map(infer, tree, is_leaf=_is_data_leaf)


def shard_spec_nothing(tree: Any) -> ShardingSpec:
    """Spec of the same structure that replicates everything."""
    return pytree.tree_map(lambda _: SpecReplicate(), tree, is_leaf=_is_data_leaf)


# ------------------------------------------------------------------------------------------------ split
def _bounds(length: int, parts: int) -> list[tuple[int, int]]:
    base, extra = divmod(length, parts)
    out, start = [], 0
    for i in range(parts):
        stop = start + base + (1 if i < extra else 0)
        out.append((start, stop))
        start = stop
    return out


def _split_leaf(leaf: Any, spec: ShardingSpecLeaf, n: int, even: bool) -> Sequence[Any]:
    if isinstance(spec, SpecReplicate):
        return (leaf,) * n
    if not isinstance(spec, SpecShard):
        raise TypeError(f"unknown sharding spec leaf: {type(spec)!r}")

    if isinstance(leaf, torch.Tensor):
        if leaf.ndim == 0:
            raise ValueError("a 0-dim tensor cannot be sharded")
        size = leaf.shape[spec.dim]
        if spec.do_stack:
            if size != n:
                raise ValueError(f"do_stack needs shape[{spec.dim}] == num_shards ({size} != {n})")
            return leaf.unbind(spec.dim)
        if even and size % n:
            raise ValueError(f"tensor of shape {tuple(leaf.shape)} is not evenly divisible into {n} shards on dim {spec.dim}")
        return torch.tensor_split(leaf, n, dim=spec.dim)

    if isinstance(leaf, list):
        if spec.dim != 0:
            raise ValueError(f"lists can only be sharded on dim 0, got {spec.dim}")
        if spec.do_stack:
            if len(leaf) != n:
                raise ValueError(f"do_stack needs len(list) == num_shards ({len(leaf)} != {n})")
            return leaf
        if even and len(leaf) % n:
            raise ValueError(f"list of length {len(leaf)} is not evenly divisible into {n} shards")
        return [leaf[a:b] for a, b in _bounds(len(leaf), n)]

    raise TypeError(f"SpecShard applied to a leaf that is neither a tensor nor a list ({type(leaf)!r})")


def shard_tree(tree: TTree, sharding_spec: ShardingSpec, num_shards: int, enforce_even_split: bool) -> tuple[TTree, ...]:
    """Split one global tree into ``num_shards`` trees of identical structure."""
    spec_leaves, structure = pytree.tree_flatten(sharding_spec)
    try:
        leaves = structure.flatten_up_to(tree)
    except (ValueError, TypeError) as exc:
        raise ValueError("Tree structure does not match sharding spec") from exc

    per_leaf = [_split_leaf(leaf, spec, num_shards, enforce_even_split) for leaf, spec in zip(leaves, spec_leaves, strict=True)]
    return tuple(structure.unflatten([pieces[i] for pieces in per_leaf]) for i in range(num_shards))


# ------------------------------------------------------------------------------------------------ merge
def _merge_leaf(pieces: Sequence[Any], spec: ShardingSpecLeaf) -> Any:
    if isinstance(spec, SpecReplicate):
        return pieces[0]
    if not isinstance(spec, SpecShard):
        raise TypeError(f"unknown sharding spec leaf: {type(spec)!r}")
    head = pieces[0]
    if isinstance(head, torch.Tensor):
        return torch.stack(list(pieces), dim=spec.dim) if spec.do_stack else torch.cat(list(pieces), dim=spec.dim)
    if spec.do_stack or isinstance(head, list):
        if spec.dim != 0:
            raise ValueError(f"lists can only be unsharded on dim 0, got {spec.dim}")
        if spec.do_stack:
            return list(pieces)
        merged: list[Any] = []
        for piece in pieces:
            merged.extend(piece)
        return merged
    raise TypeError(f"expected tensors or lists to merge, got {type(head)!r}")


def unshard_tree(sharded_trees: Sequence[TTree], sharding_spec: ShardingSpec) -> TTree:
    """Inverse of :func:`shard_tree`."""
    if not sharded_trees:
        raise ValueError("sharded_trees sequence cannot be empty")
    spec_leaves, structure = pytree.tree_flatten(sharding_spec)
    columns = []
    for idx, tree in enumerate(sharded_trees):
        try:
            columns.append(structure.flatten_up_to(tree))
        except (ValueError, TypeError) as exc:
            raise ValueError(f"Structure mismatch at shard {idx}: tree does not match sharding spec structure") from exc
    merged = [_merge_leaf([col[i] for col in columns], spec) for i, spec in enumerate(spec_leaves)]
    return structure.unflatten(merged)
