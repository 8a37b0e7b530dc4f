"""Property-based checks (hypothesis) of the pure-Python building blocks: layer distribution, dataset sharding and bucketing,
pytree sharding round trips, pipeline program validity on random shapes."""

import torch
from hypothesis import given, settings
from hypothesis import strategies as st

FAST = settings(max_examples=60, deadline=None)


@FAST
@given(layers=st.integers(1, 80), stages=st.integers(1, 12), pre=st.integers(0, 3), post=st.integers(0, 3))
def test_layer_distribution_partitions_the_layers(layers, stages, pre, post):
    from d9d_b200.pipelining.api import PipelineStageInfo, distribute_layers_for_pipeline_stage

    try:
        ranges = [distribute_layers_for_pipeline_stage(layers, pre, post, PipelineStageInfo(s, stages)) for s in range(stages)]
    except ValueError:
        return  # more (virtual) stages than the model can fill: rejected, never silently produces an empty stage
    assert ranges[0][0] == 0 and ranges[-1][1] == layers
    assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))  # contiguous, no overlap
    sizes = [hi - lo for lo, hi in ranges]
    assert all(n >= 1 for n in sizes)
    virtual = [n + (pre if i == 0 else 0) + (post if i == stages - 1 else 0) for i, n in enumerate(sizes)]
    assert max(virtual) - min(virtual) <= 1  # balanced once the embedding / head weights are counted


class _Range(torch.utils.data.Dataset):
    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return i

    def sort_key(self, i):
        return (i * 7919) % 13


@FAST
@given(n=st.integers(1, 200), shards=st.integers(1, 9), chunked=st.booleans(), pad=st.booleans())
def test_sharded_dataset_covers_everything_exactly_once(n, shards, chunked, pad):
    from d9d_b200.dataset import ShardedDataset, ShardIndexingMode

    mode = ShardIndexingMode.chunked if chunked else ShardIndexingMode.sequential
    views = [ShardedDataset(_Range(n), shards, r, mode, pad) for r in range(shards)]
    seen = [[v[i] for i in range(len(v))] for v in views]
    if pad:
        assert len({len(s) for s in seen}) == 1  # every rank runs the same number of steps
        assert set().union(*map(set, seen)) == set(range(n))
    else:
        flat = [x for s in seen for x in s]
        assert sorted(flat) == list(range(n))


@FAST
@given(n=st.integers(1, 150), buffer=st.integers(1, 40), pack=st.integers(1, 10), seed=st.integers(0, 5))
def test_buffer_sorted_dataset_is_a_windowed_permutation(n, buffer, pack, seed):
    from d9d_b200.dataset import BufferSortedDataset

    data = BufferSortedDataset(_Range(n), buffer_size=buffer, pack_size=pack, init_seed=seed)
    order = [data[i] for i in range(n)]
    assert sorted(order) == list(range(n))
    for start in range(0, n, buffer):  # shuffling never crosses a window
        assert sorted(order[start:start + buffer]) == list(range(start, min(start + buffer, n)))
    again = BufferSortedDataset(_Range(n), buffer_size=buffer, pack_size=pack, init_seed=seed)
    assert [again[i] for i in range(n)] == order  # reproducible for a seed


@FAST
@given(rows=st.integers(1, 6), shards=st.integers(1, 6), extra=st.integers(1, 4))
def test_shard_tree_round_trip(rows, shards, extra):
    from d9d_b200.core.sharding import shard_spec_on_dim, shard_tree, unshard_tree

    batch = rows * shards
    tree = {"a": torch.arange(batch * extra).view(batch, extra), "nested": {"ids": list(range(batch)), "flag": "x"},
            "tuple": (torch.arange(batch), 3.5)}
    spec = shard_spec_on_dim(tree, dim=0)
    parts = shard_tree(tree, spec, num_shards=shards, enforce_even_split=True)
    assert len(parts) == shards and all(p["a"].shape[0] == rows for p in parts)
    back = unshard_tree(parts, spec)
    assert torch.equal(back["a"], tree["a"]) and back["nested"] == tree["nested"] and torch.equal(back["tuple"][0], tree["tuple"][0])


@settings(max_examples=25, deadline=None)
@given(pp=st.integers(2, 6), microbatches=st.integers(1, 12), schedule=st.sampled_from(["bfs", "1f1b", "zb1p", "zbv", "dualpipe"]),
       per_rank=st.integers(1, 3))
def test_random_pipeline_programs_are_complete_and_deadlock_free(pp, microbatches, schedule, per_rank):
    from d9d_b200.pipelining.infra.communications import validate_program
    from d9d_b200.pipelining.infra.programs import (DualPipeVPipelineProgramBuilder, Interleaved1F1BPipelineProgramBuilder,
                                                    LoopedBFSPipelineProgramBuilder, ZeroBubbleVPipelineProgramBuilder)
    from d9d_b200.pipelining.infra.topology import build_stage_to_host_rank_topology

    builder = {"bfs": lambda: LoopedBFSPipelineProgramBuilder(per_rank),
               "1f1b": lambda: Interleaved1F1BPipelineProgramBuilder(per_rank, enable_zero_bubble=False),
               "zb1p": lambda: Interleaved1F1BPipelineProgramBuilder(per_rank, enable_zero_bubble=True),
               "zbv": ZeroBubbleVPipelineProgramBuilder, "dualpipe": DualPipeVPipelineProgramBuilder}[schedule]()
    try:
        program = builder.compose(num_microbatches=microbatches, pp_size=pp)
    except ValueError:
        return  # some schedules need a minimum number of microbatches
    num_stages = builder.num_stages_per_rank * pp
    topology = build_stage_to_host_rank_topology(num_stages=num_stages, pp_size=pp, style=builder.topology_style)
    validate_program(program, topology, num_stages, microbatches, has_backward=True)
