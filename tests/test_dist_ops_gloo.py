"""Convenience collectives (fixed / variadic shapes, objects) and distributed metric synchronisation on gloo."""

import pytest
import torch

from tests.dist_utils import run_distributed

pytestmark = pytest.mark.dist


def _worker(rank, world_size):
    import torch.distributed as dist

    from d9d_b200.core.dist_ops import (all_gather, all_gather_object, all_gather_variadic_shape, gather, gather_object,
                                        gather_variadic_shape)
    from d9d_b200.metric.impl.aggregation import WeightedMeanMetric

    dist.init_process_group("gloo")  # env:// rendezvous prepared by run_distributed
    group = dist.distributed_c10d._get_default_group()
    x = torch.full((2, 3), float(rank))
    assert [t[0, 0].item() for t in all_gather(x, group)] == [0.0, 1.0, 2.0][:world_size]
    got = gather(x, group, group_dst=0)
    assert (got is None) == (rank != 0)

    ragged = torch.arange(rank + 2, dtype=torch.float32).view(1, rank + 2) + 10 * rank  # different shape per rank
    parts = all_gather_variadic_shape(ragged, group)
    assert [tuple(p.shape) for p in parts] == [(1, r + 2) for r in range(world_size)]
    assert all(torch.equal(p, torch.arange(r + 2, dtype=torch.float32).view(1, r + 2) + 10 * r) for r, p in enumerate(parts))
    on_dst = gather_variadic_shape(ragged, group, group_dst=1)
    if rank == 1:
        assert [tuple(p.shape) for p in on_dst] == [(1, r + 2) for r in range(world_size)]
    else:
        assert on_dst is None

    objs = all_gather_object({"rank": rank, "tag": "x" * rank}, group)
    assert [o["rank"] for o in objs] == list(range(world_size))
    gathered = gather_object(rank, group, group_dst=0)
    assert gathered == (list(range(world_size)) if rank == 0 else None)

    metric = WeightedMeanMetric()
    metric.update(torch.tensor(float(rank + 1)), torch.tensor(float(rank + 1)))

    class Ctx:  # the metric only needs to know that the job is distributed
        is_distributed = True

    metric.sync(Ctx())
    want = sum((r + 1) ** 2 for r in range(world_size)) / sum(r + 1 for r in range(world_size))
    assert abs(float(metric.compute()) - want) < 1e-6


def test_dist_ops_and_metric_sync():
    run_distributed(_worker, 3)


def _metric_checkpoint_worker(rank, world_size, tmp):
    import torch.distributed as dist
    import torch.distributed.checkpoint as dcp

    from d9d_b200.metric.impl.aggregation import SumMetric, WeightedMeanMetric
    from d9d_b200.metric.impl.container import ComposeMetric

    dist.init_process_group("gloo")

    def fresh():
        return ComposeMetric({"tokens": SumMetric(), "loss": WeightedMeanMetric()})

    metrics = fresh()
    metrics["tokens"].update(torch.tensor(10.0 * (rank + 1)))  # rank-specific partial sums in the middle of a logging period
    metrics["loss"].update(torch.tensor([float(rank)]), torch.tensor([1.0 + rank]))
    dcp.save({"metrics": metrics}, checkpoint_id=tmp)

    restored = fresh()
    dcp.load({"metrics": restored}, checkpoint_id=tmp)
    restored["tokens"].update(torch.tensor(1.0))  # accumulation continues after the restart
    restored.sync(None)
    values = restored.compute()
    assert float(values["tokens"]) == 10.0 * sum(r + 1 for r in range(world_size)) + world_size
    expected_loss = sum(r * (1.0 + r) for r in range(world_size)) / sum(1.0 + r for r in range(world_size))
    assert abs(float(values["loss"]) - expected_loss) < 1e-6


def test_metric_state_survives_a_distributed_checkpoint(tmp_path):
    """Rank-specific partial metric sums are checkpointed under rank-independent keys: the global totals must come back."""
    run_distributed(_metric_checkpoint_worker, 2, str(tmp_path / "ckpt"))
