"""Column / row tensor parallelism (with and without sequence parallelism) on CPU/gloo against a single-process MLP."""

import pytest
import torch
from torch import nn

from tests.dist_utils import run_distributed

pytestmark = pytest.mark.dist


class _MLP(nn.Module):
    def __init__(self):
        super().__init__()
        from d9d_b200.module.block.linear import Linear

        self.up = Linear(32, 64, bias=True)
        self.down = Linear(64, 32, bias=True)

    def forward(self, x):
        return self.down(torch.nn.functional.gelu(self.up(x)))


def _worker(rank, world_size, sequence_parallel):
    from torch.distributed.device_mesh import init_device_mesh
    from torch.distributed.tensor import DTensor

    from d9d_b200.module.parallelism.api import parallelize_colwise, parallelize_rowwise

    mesh = init_device_mesh("cpu", (world_size,), mesh_dim_names=("tp",))
    torch.manual_seed(0)
    ref = _MLP()
    tp = _MLP()
    tp.load_state_dict(ref.state_dict())
    parallelize_colwise(tp.up, mesh, sequence_parallel=sequence_parallel)
    parallelize_rowwise(tp.down, mesh, sequence_parallel=sequence_parallel)
    assert isinstance(tp.up.weight.data, DTensor) and tp.up.weight.to_local().shape == (64 // world_size, 32)
    assert tp.down.weight.to_local().shape == (32, 64 // world_size)

    x = torch.randn(2, 8, 32)
    x_ref = x.clone().requires_grad_()
    y_ref = ref(x_ref)
    y_ref.square().sum().backward()

    if sequence_parallel:
        x_in = x.chunk(world_size, dim=1)[rank].clone().requires_grad_()
    else:
        x_in = x.clone().requires_grad_()
    y = tp(x_in)
    want_y = y_ref.chunk(world_size, dim=1)[rank] if sequence_parallel else y_ref
    torch.testing.assert_close(y, want_y.detach(), rtol=1e-4, atol=1e-5)
    # each rank back-propagates its share of the loss; with replicated outputs every rank holds the whole loss
    (y.square().sum()).backward()
    want_dx = x_ref.grad.chunk(world_size, dim=1)[rank] if sequence_parallel else x_ref.grad
    torch.testing.assert_close(x_in.grad, want_dx, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(tp.up.weight.grad.to_local(), ref.up.weight.grad.chunk(world_size, dim=0)[rank], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(tp.down.weight.grad.to_local(), ref.down.weight.grad.chunk(world_size, dim=1)[rank], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(tp.up.bias.grad.to_local(), ref.up.bias.grad.chunk(world_size, dim=0)[rank], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("sequence_parallel", [False, True])
def test_colwise_rowwise_mlp(sequence_parallel):
    run_distributed(_worker, 2, sequence_parallel)
