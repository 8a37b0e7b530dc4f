"""FSDP with its all-gather / reduce-scatter replaced by the peer-memory pull protocol (``_peer_memory_fsdp.py``), on gloo
with an emulated symmetric arena: every rank's staging buffer is visible to its peers through a snapshot all-gather."""

import pytest
import torch
from torch import nn

from tests.dist_utils import run_distributed

pytestmark = pytest.mark.dist


class _EmulatedArena:
    """Stand-in for ``SymmetricArena`` on CPU: ``snapshot()`` returns every rank's buffer (taken after the barrier)."""

    calls = 0

    def __init__(self, nbytes, device, group):
        self.buffer = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        self.group = group

    def barrier(self):
        import torch.distributed as dist

        dist.barrier(group=self.group)

    def snapshot(self):
        import torch.distributed as dist

        type(self).calls += 1
        out = [torch.empty_like(self.buffer) for _ in range(self.group.size())]
        dist.all_gather(out, self.buffer, group=self.group)
        return out


def _model():
    torch.manual_seed(0)
    return nn.Sequential(nn.Linear(8, 16), nn.GELU(), nn.Linear(16, 6), nn.GELU(), nn.Linear(6, 4))


def _train(model, xs, rank, steps=3):
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    for s in range(steps):
        model(xs[s][rank]).square().sum().backward()
        opt.step()
        opt.zero_grad()


def _arena_factory(nbytes, device, group):
    return _EmulatedArena(nbytes, device, group)


def _worker(rank, world_size):
    from torch.distributed.device_mesh import init_device_mesh
    from torch.distributed.tensor import DTensor

    from d9d_b200.module.parallelism.api import parallelize_fsdp

    mesh = init_device_mesh("cpu", (world_size,), mesh_dim_names=("dp_cp_shard",))
    peer, plain = _model(), _model()
    for unit in (peer[0], peer[2], peer):  # two FSDP units + the root, all sharing the staging buffers of the group
        parallelize_fsdp(unit, mesh, peer_memory_arena_factory=_arena_factory)
    for unit in (plain[0], plain[2], plain):
        parallelize_fsdp(unit, mesh)

    xs = [[torch.randn(3, 8, generator=torch.Generator().manual_seed(100 * s + r)) for r in range(world_size)] for s in range(3)]
    _train(peer, xs, rank)
    _train(plain, xs, rank)
    assert _EmulatedArena.calls > 0  # the custom collectives really ran
    for (name, a), (_, b) in zip(peer.named_parameters(), plain.named_parameters(), strict=True):
        assert isinstance(a.data, DTensor) and a.to_local().shape == b.to_local().shape
        torch.testing.assert_close(a.to_local(), b.to_local(), rtol=1e-5, atol=1e-6, msg=lambda m, name=name: f"{name}: {m}")


@pytest.mark.parametrize("world", [2, 3])
def test_fsdp_with_peer_memory_collectives_matches_nccl_style_collectives(world):
    run_distributed(_worker, world)
