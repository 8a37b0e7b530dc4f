"""Mapper algebra, optimizer convergence, mesh domains, seeds, garbage collector, profiler, metric collector."""

import copy
import gc
import tarfile

import pytest
import torch

from tests.dist_utils import run_distributed


# ------------------------------------------------------------------------------------------------ mapper algebra
def _groups(m):
    return {(tuple(sorted(g.inputs)), tuple(sorted(g.outputs))) for g in m.state_dependency_groups()}


def test_mapper_composition_rules():
    from d9d_b200.model_state.mapper.compose import (ModelStateMapperParallel, ModelStateMapperPrefixScope, ModelStateMapperSequential,
                                                     ModelStateMapperShard)
    from d9d_b200.model_state.mapper.leaf import (ModelStateMapperChunkTensors, ModelStateMapperIdentity, ModelStateMapperRename,
                                                  ModelStateMapperSelectChildModules, ModelStateMapperSqueeze, ModelStateMapperStackTensors,
                                                  ModelStateMapperUnsqueeze)

    with pytest.raises(ValueError):  # two mappers may not consume / produce the same key
        ModelStateMapperParallel([ModelStateMapperIdentity("a"), ModelStateMapperRename("a", "b")])
    with pytest.raises(ValueError):
        ModelStateMapperParallel([ModelStateMapperRename("a", "c"), ModelStateMapperRename("b", "c")])

    # sequential: keys a stage does not touch are carried through; groups that meet downstream merge
    seq = ModelStateMapperSequential([
        ModelStateMapperParallel([ModelStateMapperRename("x0", "y0"), ModelStateMapperRename("x1", "y1")]),
        ModelStateMapperStackTensors(["y0", "y1"], "stacked", dim=0),
    ])
    assert _groups(seq) == {(("x0", "x1"), ("stacked",))}
    out = seq.apply({"x0": torch.ones(2), "x1": torch.zeros(2)})
    assert out["stacked"].shape == (2, 2)
    carry = ModelStateMapperSequential([ModelStateMapperParallel([ModelStateMapperRename("a", "b"), ModelStateMapperIdentity("k")]),
                                        ModelStateMapperRename("b", "c")])
    assert _groups(carry) == {(("a",), ("c",)), (("k",), ("k",))}

    scoped = ModelStateMapperPrefixScope(ModelStateMapperChunkTensors("w", ["w0", "w1"], dim=0), source_prefix="hf.", target_prefix="native.")
    assert _groups(scoped) == {(("hf.w",), ("native.w0", "native.w1"))}
    res = scoped.apply({"hf.w": torch.arange(4.0)})
    assert res["native.w1"].tolist() == [2.0, 3.0]

    many = ModelStateMapperParallel([ModelStateMapperIdentity(f"p{i}") for i in range(5)])
    shards = [ModelStateMapperShard(many, total_shards=2, current_shard=i) for i in range(2)]
    seen = [g for s in shards for g in _groups(s)]
    assert len(seen) == 5 and len(set(seen)) == 5  # a partition of the groups

    child = ModelStateMapperSelectChildModules(["weight", "bias"], parent_name="layer")
    assert child.apply({"layer.weight": torch.ones(1)})["weight"].item() == 1.0
    assert ModelStateMapperUnsqueeze("t", 0).apply({"t": torch.ones(3)})["t"].shape == (1, 3)
    assert ModelStateMapperSqueeze("t", 0).apply({"t": torch.ones(1, 3)})["t"].shape == (3,)


# ------------------------------------------------------------------------------------------------ optimizer
@pytest.mark.parametrize("state_dtype", [torch.float32, torch.bfloat16])
def test_stochastic_adamw_converges_on_cpu(state_dtype):
    from d9d_b200.optim.stochastic import StochasticAdamW

    torch.manual_seed(0)
    target = torch.randn(64)
    p = torch.nn.Parameter(torch.zeros(64, dtype=torch.bfloat16))
    opt = StochasticAdamW([p], lr=5e-2, weight_decay=0.0, state_dtype=state_dtype)
    for _ in range(400):
        loss = (p.float() - target).square().mean()
        loss.backward()
        opt.step()
        opt.zero_grad()
    assert float((p.float() - target).square().mean()) < 1e-3
    state = copy.deepcopy(opt.state_dict())  # load_state_dict re-attaches tensors without copying them
    assert "_d9d_generator_state" in state  # the rounding stream is checkpointed
    clone = StochasticAdamW([torch.nn.Parameter(p.detach().clone())], lr=5e-2, weight_decay=0.0, state_dtype=state_dtype)
    clone.load_state_dict(state)
    g = torch.randn(64)
    for o in (opt, clone):
        o.param_groups[0]["params"][0].grad = g.clone().bfloat16()
        o.step()
    assert torch.equal(opt.param_groups[0]["params"][0], clone.param_groups[0]["params"][0])  # bit-exact continuation


# ------------------------------------------------------------------------------------------------ mesh domains / seeds
def _mesh_worker(rank, world_size):
    import random

    from d9d_b200.core.dist_context import BATCH_DOMAIN, DENSE_DOMAIN, EXPERT_DOMAIN, FLAT_DOMAIN, REGULAR_DOMAIN, DeviceMeshParameters
    from d9d_b200.internals.determinism import set_seeds

    ctx = DeviceMeshParameters(pipeline_parallel=2, data_parallel_replicate=2, data_parallel_shard=2, expert_parallel=2).build()
    assert dict(zip(ctx.mesh_for(REGULAR_DOMAIN).mesh_dim_names, ctx.mesh_for(REGULAR_DOMAIN).shape)) == {
        "pp": 2, "dp_replicate": 2, "dp_shard": 2, "cp_shard": 1, "cp_replicate": 1, "tp": 1}
    assert ctx.mesh_for(DENSE_DOMAIN)["dp_cp_shard"].size() == 2
    expert = ctx.mesh_for(EXPERT_DOMAIN)
    assert expert["ep_shard"].size() == 2 and expert["ep_replicate"].size() == 2
    assert ctx.mesh_for(BATCH_DOMAIN)["dp"].size() == 4 and ctx.mesh_for(FLAT_DOMAIN).size() == 8
    with pytest.raises(ValueError):
        ctx.mesh_for("nope")

    set_seeds(ctx, seed=100)
    draw = torch.tensor([random.random(), float(torch.rand(()))])
    everyone = [torch.zeros(2) for _ in range(world_size)]
    torch.distributed.all_gather(everyone, draw)
    pp = ctx.mesh_for(REGULAR_DOMAIN).get_local_rank("pp")
    pps = [torch.zeros(1) for _ in range(world_size)]
    torch.distributed.all_gather(pps, torch.tensor([float(pp)]))
    for other, other_pp in zip(everyone, pps):  # same pipeline stage <=> same random streams
        assert torch.equal(other, draw) == (int(other_pp) == pp)


@pytest.mark.dist
def test_mesh_domains_and_seeds():
    run_distributed(_mesh_worker, 8)


# ------------------------------------------------------------------------------------------------ small components
class _Ctx:
    is_distributed = False
    is_main_process = True
    is_local_main_process = True
    global_rank = 0

    class _Log:
        def info(self, *_a, **_k): ...
        def debug(self, *_a, **_k): ...

    logger = _Log()

    class _Params:
        is_distributed = False

    mesh_params = _Params()
    current_device = torch.device("cpu")

    def wait_world(self): ...


def test_manual_garbage_collector_controls_the_interpreter_gc():
    from d9d_b200.loop.component import ManualGarbageCollector, Stepper
    from d9d_b200.loop.config import GarbageCollectionConfig

    stepper = Stepper(0, 10)
    assert gc.isenabled()
    with ManualGarbageCollector(_Ctx(), GarbageCollectionConfig(period_steps=3), stepper) as collector:
        assert not gc.isenabled()  # automatic collection is off inside the loop (no random pauses on some ranks)
        for _ in range(6):
            collector.collect_periodic()
            stepper.step()
        collector.collect_forced()
    assert gc.isenabled()


def test_profiler_writes_one_compressed_trace_per_cycle(tmp_path):
    from d9d_b200.internals.profiling import Profiler

    prof = Profiler(save_dir=tmp_path, period_steps=3, warmup_steps=1, active_steps=1, dist_context=_Ctx())
    with prof.open(start_step=0) as p:
        for _ in range(6):
            torch.ones(8).sum()
            p.step()
    traces = sorted(tmp_path.rglob("*.tar.gz"))
    assert len(traces) == 2
    with tarfile.open(traces[0]) as tar:
        assert any(name.endswith(".json") for name in tar.getnames())


def test_async_metric_collector_round_trip():
    from d9d_b200.internals.metric_collector import AsyncMetricCollector
    from d9d_b200.metric.impl.aggregation import SumMetric
    from d9d_b200.metric.impl.container import ComposeMetric

    metrics = ComposeMetric({"n": SumMetric()})
    collector = AsyncMetricCollector(metrics)
    collector.bind("cpu")
    metrics.children["n"].update(torch.tensor([2.0, 3.0]))
    collector.schedule_collection(_Ctx())
    assert collector.collect_results() == {"n": 5.0}
    assert float(metrics.compute()["n"]) == 0.0  # collecting resets the metric
    with pytest.raises(Exception):
        collector.collect_results()  # nothing scheduled
    collector.unbind()
