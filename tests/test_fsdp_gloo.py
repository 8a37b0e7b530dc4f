"""FSDP / HSDP hooks on CPU/gloo: sharded parameters, gradients equal to the single-process SUM over ranks."""

import pytest
import torch
from torch import nn

from tests.dist_utils import run_distributed

pytestmark = pytest.mark.dist


def _model():
    torch.manual_seed(0)
    return nn.Sequential(nn.Linear(8, 16), nn.GELU(), nn.Linear(16, 4))


def _worker(rank, world_size, hybrid):
    from torch.distributed.device_mesh import init_device_mesh
    from torch.distributed.tensor import DTensor

    from d9d_b200.module.parallelism.api import parallelize_fsdp, parallelize_hsdp

    model, ref = _model(), _model()
    if hybrid:
        mesh = init_device_mesh("cpu", (2, world_size // 2), mesh_dim_names=("dp_replicate", "dp_cp_shard"))
        parallelize_hsdp(model, mesh, shard_dim="dp_cp_shard")
    else:
        mesh = init_device_mesh("cpu", (world_size,), mesh_dim_names=("dp_cp_shard",))
        parallelize_fsdp(model, mesh)
    params = dict(model.named_parameters())
    assert all(isinstance(p.data, DTensor) for p in params.values())
    assert params["0.weight"].to_local().shape[0] < 16  # really sharded

    xs = [torch.randn(3, 8, generator=torch.Generator().manual_seed(10 + r)) for r in range(world_size)]
    model(xs[rank]).square().sum().backward()
    # FSDP reduce-scatters (SUM) inside the shard group only; summing over the replicate dim is the gradient
    # synchroniser's job, so right after backward a rank holds the sum over *its shard group*
    shard = world_size // 2 if hybrid else world_size
    group_ranks = range(rank // shard * shard, rank // shard * shard + shard)
    for r in group_ranks:
        ref(xs[r]).square().sum().backward()
    for name, p in params.items():
        got = p.grad.full_tensor() if isinstance(p.grad, DTensor) else p.grad
        torch.testing.assert_close(got, dict(ref.named_parameters())[name].grad, rtol=1e-4, atol=1e-5, msg=lambda m: f"{name}: {m}")  # noqa: B023
    if hybrid:  # the synchroniser then completes the reduction over the replicas
        from d9d_b200.internals.grad_sync.synchronizer import find_reduce_mesh

        assert find_reduce_mesh(params["0.weight"].data).mesh_dim_names == ("dp_replicate",)


@pytest.mark.parametrize("hybrid", [False, True])
def test_fsdp_and_hsdp_sum_gradients(hybrid):
    run_distributed(_worker, 4 if hybrid else 2, hybrid)
