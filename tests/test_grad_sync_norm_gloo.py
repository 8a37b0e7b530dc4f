"""Synthetic gradient synchronisation / distributed norm checks on a 2 x 2 gloo mesh.

Follows the reference's ``internals/grad_sync/test_e2e.py`` and ``internals/grad_norm/test_{correctness,grouping}.py``:
hand-made DTensor parameters with Shard / Replicate mixes and (param, grad) dtype pairs, exact expected sums,
"not ready" detection, hook clean-up, bucket splitting, and norms checked against a clip over the gathered gradients.
"""

import math

import pytest
import torch
from torch import nn

from tests.dist_utils import run_distributed

pytestmark = pytest.mark.dist


def _make_params(mesh, param_dtype, grad_dtype):
    from torch.distributed.tensor import DTensor, Replicate, Shard

    def with_grad_dtype(p):
        if grad_dtype != param_dtype:
            p.grad_dtype = grad_dtype
        return p

    def dt(local, placements):
        return with_grad_dtype(nn.Parameter(DTensor.from_local(local.to(param_dtype), mesh, placements, run_check=False)))

    return {
        "replicated": dt(torch.zeros(6, 4), (Replicate(), Replicate())),  # summed over all four ranks
        "hsdp": dt(torch.zeros(3, 4), (Replicate(), Shard(0))),  # summed over the replicate dim only
        "sharded": dt(torch.zeros(2, 5), (Shard(0), Shard(1))),  # nothing to reduce
        "plain": with_grad_dtype(nn.Parameter(torch.zeros(7, dtype=param_dtype))),  # not a DTensor: local accumulation only
    }


def _backward_rank_valued(params, value: float):
    """Every element of every gradient receives ``value`` (through a real autograd pass, so the hooks fire)."""
    from torch.distributed.tensor import DTensor

    loss = 0
    for p in params.values():
        local = p.to_local() if isinstance(p, DTensor) else p
        loss = loss + (local.float() * value).sum()
    loss.backward()


def _sync_worker(rank, world_size, dtype_pairs):
    import torch.distributed as dist
    from torch.distributed.device_mesh import init_device_mesh

    dist.init_process_group("gloo")
    mesh = init_device_mesh("cpu", (2, 2), mesh_dim_names=("dp_replicate", "dp_shard"))
    for param_dtype, grad_dtype in dtype_pairs:
        try:
            _check_sync(rank, mesh, param_dtype, grad_dtype)
        except AssertionError as exc:
            raise AssertionError(f"[param {param_dtype}, grad {grad_dtype}] {exc}") from exc


def _check_sync(rank, mesh, param_dtype, grad_dtype):
    from torch.distributed.tensor import DTensor

    from d9d_b200.internals.grad_sync import GradientSynchronizer

    params = _make_params(mesh, param_dtype, grad_dtype)
    plist = list(params.values())
    sync = GradientSynchronizer([plist], bucket_size_mb=1, require_accumulations=2)
    sync.bind()

    def local_grad(p):
        return p.grad.to_local() if isinstance(p.grad, DTensor) else p.grad

    for p in plist:  # gradients exist up front, are zero, have the requested dtype and alias the flat arenas
        assert local_grad(p).dtype == grad_dtype and float(local_grad(p).abs().sum()) == 0
    arena_ptrs = {(a.buffer.data_ptr(), a.buffer.data_ptr() + a.buffer.numel() * a.buffer.element_size()) for a in sync.arenas}
    for p in plist:
        ptr = local_grad(p).data_ptr()
        assert any(lo <= ptr < hi for lo, hi in arena_ptrs)

    _backward_rank_valued(params, float(rank + 1))
    with pytest.raises(ValueError):
        sync.wait()  # one of two accumulations: the buckets are not ready
    _backward_rank_valued(params, float(rank + 1))
    sync.wait()

    replicate_coord, shard_coord = mesh.get_coordinate()
    everyone = 2.0 * sum(r + 1 for r in range(4))  # two rounds, ranks contribute 1..4
    # ranks are laid out row-major: rank = replicate * 2 + shard; the hsdp parameter sums over the replicate dim
    same_shard = 2.0 * sum(r * 2 + shard_coord + 1 for r in range(2))
    own = 2.0 * (rank + 1)
    assert torch.all(local_grad(params["replicated"]) == everyone)
    assert torch.all(local_grad(params["hsdp"]) == same_shard)
    assert torch.all(local_grad(params["sharded"]) == own)
    assert torch.all(local_grad(params["plain"]) == own)
    del replicate_coord

    sync.zero_grad()
    assert all(float(local_grad(p).abs().sum()) == 0 for p in plist)
    sync.unbind()
    assert all(p.grad is None for p in plist)
    _backward_rank_valued(params, 1.0)  # hooks are gone: a plain backward works and nothing is reduced
    assert torch.all(local_grad(params["replicated"]) == 1.0)


def test_gradient_synchronizer_sums_exactly():
    run_distributed(_sync_worker, 4, [(torch.float32, torch.float32), (torch.bfloat16, torch.float32), (torch.bfloat16, torch.bfloat16)])


def test_bucket_splitting_respects_the_byte_budget():
    from d9d_b200.internals.grad_sync import GradientSynchronizer

    params = [nn.Parameter(torch.zeros(300_000)) for _ in range(5)]  # 1.2 MB each
    sync = GradientSynchronizer([params], bucket_size_mb=2, require_accumulations=1)
    sync.bind()
    try:
        assert len(sync.arenas) == 1  # one flat buffer for the class ...
        assert len(sync.buckets) == 5  # ... cut into buckets that stay below 2 MiB (1.2 + 1.2 would exceed it)
    finally:
        sync.unbind()
    sync = GradientSynchronizer([params], bucket_size_mb=3, require_accumulations=1)
    sync.bind()
    try:
        assert len(sync.buckets) == 3  # 2 + 2 + 1
    finally:
        sync.unbind()


# ------------------------------------------------------------------------------------------------ norms
def _norm_worker(rank, world_size, norm_types):
    import torch.distributed as dist
    from torch.distributed.device_mesh import init_device_mesh

    dist.init_process_group("gloo")
    mesh = init_device_mesh("cpu", (2, 2), mesh_dim_names=("dp_replicate", "dp_shard"))
    for norm_type in norm_types:
        try:
            _check_norm(mesh, norm_type)
        except AssertionError as exc:
            raise AssertionError(f"[norm {norm_type}] {exc}") from exc


def _check_norm(mesh, norm_type):
    from torch.distributed.tensor import DTensor, Replicate, Shard

    from d9d_b200.internals.grad_norm import clip_grad_norm_distributed_, group_parameters_for_norm

    torch.manual_seed(0)
    full = {"a": torch.randn(8, 6), "b": torch.randn(4, 4), "c": torch.randn(10)}

    def param(name, placements):
        from torch.distributed.tensor import distribute_tensor

        p = nn.Parameter(distribute_tensor(torch.zeros_like(full[name]), mesh, placements))
        p.grad = distribute_tensor(full[name].clone(), mesh, placements)
        return p

    a = param("a", (Replicate(), Shard(0)))
    b = param("b", (Replicate(), Replicate()))
    c = nn.Parameter(torch.zeros(10))
    c.grad = full["c"].clone()
    frozen = nn.Parameter(torch.zeros(3), requires_grad=False)

    groups = group_parameters_for_norm([a, b, c, frozen])
    keys = list(groups)
    assert sum(len(v) for v in groups.values()) == 3  # the frozen parameter is not part of any group
    assert keys[0].shard_meshes is not None and all(k.shard_meshes is None for k in keys[1:])  # sharded groups lead
    assert groups[keys[0]] == [a]

    expected = torch.nn.utils.get_total_norm(list(full.values()), norm_type=norm_type)
    total = clip_grad_norm_distributed_(groups, max_norm=None, norm_type=norm_type, pp_mesh=None)
    torch.testing.assert_close(total, expected, rtol=1e-5, atol=1e-6)

    limit = 0.25 * float(expected)
    clip_grad_norm_distributed_(groups, max_norm=limit, norm_type=norm_type, pp_mesh=None)
    scale = min(1.0, limit / (float(expected) + 1e-6))
    torch.testing.assert_close(a.grad.full_tensor(), full["a"] * scale, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(b.grad.full_tensor(), full["b"] * scale, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(c.grad, full["c"] * scale, rtol=1e-5, atol=1e-6)
    assert isinstance(a.grad, DTensor)

    # a pending (not yet applied) scale is folded into the reported norm and absorbs the clip coefficient
    pending = torch.tensor([0.5])
    before = a.grad.to_local().clone()
    total = clip_grad_norm_distributed_(groups, max_norm=1e-3, norm_type=norm_type, pp_mesh=None, pending_scale=pending)
    torch.testing.assert_close(total, expected * scale * 0.5, rtol=1e-4, atol=1e-6)
    assert torch.equal(a.grad.to_local(), before)  # gradients untouched ...
    assert math.isclose(float(pending) * float(expected * scale), 1e-3, rel_tol=1e-3)  # ... the scalar carries the clip


def test_distributed_norm_matches_gathered_clip():
    run_distributed(_norm_worker, 4, [2.0, 1.0, math.inf])
