"""Sharding trees, LR schedules, datasets, padding/pooling, events, stepper."""

import math

import pytest
import torch

from d9d_b200.core.sharding import SpecReplicate, SpecShard, shard_spec_on_dim, shard_tree, unshard_tree


def test_shard_tree_round_trip():
    tree = {"x": torch.arange(24).view(6, 4), "meta": {"ids": [1, 2, 3, 4, 5, 6], "flag": "keep"}, "scalar": torch.tensor(3.0)}
    spec = {"x": SpecShard(0), "meta": {"ids": SpecShard(0), "flag": SpecReplicate()}, "scalar": SpecReplicate()}
    shards = shard_tree(tree, spec, num_shards=3, enforce_even_split=True)
    assert len(shards) == 3 and shards[1]["x"].tolist() == [[8, 9, 10, 11], [12, 13, 14, 15]]
    assert shards[2]["meta"]["ids"] == [5, 6] and shards[0]["meta"]["flag"] == "keep"
    back = unshard_tree(shards, spec)
    assert torch.equal(back["x"], tree["x"]) and back["meta"]["ids"] == tree["meta"]["ids"]


def test_shard_tree_stack_mode_and_auto_spec():
    tree = {"loss": torch.arange(4.0), "hidden": torch.randn(8, 3)}
    spec = shard_spec_on_dim(tree, dim=0)
    shards = shard_tree(tree, spec, num_shards=4, enforce_even_split=True)
    assert shards[0]["hidden"].shape == (2, 3)
    stacked = {"loss": SpecShard(0, do_stack=True)}
    per_mb = [{"loss": torch.tensor(float(i))} for i in range(3)]
    assert unshard_tree(per_mb, stacked)["loss"].tolist() == [0.0, 1.0, 2.0]
    with pytest.raises(ValueError):
        shard_tree({"x": torch.arange(5)}, {"x": SpecShard(0)}, num_shards=2, enforce_even_split=True)


def test_piecewise_schedule_builder_and_config():
    from d9d_b200.lr_scheduler.piecewise import (CurveCosine, CurveLinear, PiecewiseSchedulerConfig, piecewise_schedule,
                                                 piecewise_scheduler_from_config)
    from d9d_b200.lr_scheduler.visualizer import lr_history

    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=2.0)
    sched = (piecewise_schedule(0.0, total_steps=100).until_percentage(0.1, 1.0, CurveLinear())
             .for_steps(40, 1.0, CurveLinear()).fill_rest(0.1, CurveCosine()).build(opt))
    lrs = []
    for _ in range(100):
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        sched.step()
    assert lrs[0] == 0.0 and abs(lrs[5] - 1.0) < 1e-6 and abs(lrs[10] - 2.0) < 1e-6 and abs(lrs[49] - 2.0) < 1e-6
    assert lrs[99] < lrs[60] < 2.0 and lrs[99] >= 0.2 - 1e-6
    mid = 0.1 + (1.0 - 0.1) * 0.5 * (1 + math.cos(math.pi * 25 / 50))
    assert abs(lrs[75] - 2.0 * mid) < 1e-6

    cfg = PiecewiseSchedulerConfig.model_validate({"initial_multiplier": 0.0, "phases": [
        {"mode": "steps", "steps": 10, "target_multiplier": 1.0, "curve": {"type": "linear"}},
        {"mode": "rest", "target_multiplier": 0.0, "curve": {"type": "poly", "power": 2.0}}]})
    hist = lr_history(lambda o: piecewise_scheduler_from_config(cfg, o, total_steps=20), num_steps=20, init_lr=1.0)
    assert hist[0] == 0.0 and abs(hist[10] - 1.0) < 1e-6 and abs(hist[19] - 0.19) < 1e-6
    with pytest.raises(ValueError):
        piecewise_schedule(0.0).until_percentage(0.5, 1.0, CurveLinear())


def test_sharded_and_buffer_sorted_datasets():
    from d9d_b200.dataset import BufferSortedDataset, ShardedDataset
    from d9d_b200.dataset.sharded import ShardIndexingMode

    base = list(range(10))
    seq = [ShardedDataset(base, 3, r, ShardIndexingMode.sequential, False) for r in range(3)]
    assert [list(s[i] for i in range(len(s))) for s in seq] == [[0, 3, 6, 9], [1, 4, 7], [2, 5, 8]]
    chunk = [ShardedDataset(base, 3, r, ShardIndexingMode.chunked, True) for r in range(3)]
    assert all(len(c) == 4 for c in chunk) and [chunk[2][i] for i in range(4)] == [8, 9, 9, 9]  # padded tail repeats

    class Lens(torch.utils.data.Dataset):
        def __init__(self):
            g = torch.Generator().manual_seed(0)
            self.l = torch.randint(1, 100, (64,), generator=g).tolist()
        def __len__(self): return len(self.l)
        def sort_key(self, i): return self.l[i]
        def __getitem__(self, i): return self.l[i]

    ds = BufferSortedDataset(Lens(), buffer_size=16, pack_size=4, init_seed=1)
    out = [ds[i] for i in range(len(ds))]
    assert sorted(out) == sorted(Lens().l)  # a permutation
    packs = [out[i:i + 4] for i in range(0, 64, 4)]
    spread = sum(max(p) - min(p) for p in packs) / len(packs)
    base_spread = sum(max(p) - min(p) for p in [Lens().l[i:i + 4] for i in range(0, 64, 4)]) / 16
    assert spread < 0.5 * base_spread  # packs hold similar lengths
    # resume mid-epoch: the checkpoint carries the RNG and the current window
    ds_a = BufferSortedDataset(Lens(), buffer_size=16, pack_size=4, init_seed=1)
    head = [ds_a[i] for i in range(21)]
    state = ds_a.state_dict()
    tail = [ds_a[i] for i in range(21, 64)]
    ds_b = BufferSortedDataset(Lens(), buffer_size=16, pack_size=4, init_seed=999)
    ds_b.load_state_dict(state)
    assert [ds_b[i] for i in range(21, 64)] == tail and head + tail == out


def test_padding_and_pooling():
    from d9d_b200.dataset import pad_stack_1d, token_pooling_mask_from_attention_mask
    from d9d_b200.dataset.padding import PaddingSide1D
    from d9d_b200.dataset.pooling import TokenPoolingType

    items = [torch.tensor([1, 2, 3]), torch.tensor([4])]
    assert pad_stack_1d(items, pad_value=0).tolist() == [[1, 2, 3], [4, 0, 0]]
    assert pad_stack_1d(items, pad_value=-1, padding_side=PaddingSide1D.left).tolist() == [[1, 2, 3], [-1, -1, 4]]
    assert pad_stack_1d(items, pad_value=0, pad_to_multiple_of=4).shape == (2, 4)
    mask = torch.tensor([[1, 1, 1, 0], [0, 1, 1, 1]])
    assert token_pooling_mask_from_attention_mask(mask, TokenPoolingType.last).tolist() == [[0, 0, 1, 0], [0, 0, 0, 1]]
    assert token_pooling_mask_from_attention_mask(mask, TokenPoolingType.first).tolist() == [[1, 0, 0, 0], [0, 1, 0, 0]]
    assert token_pooling_mask_from_attention_mask(mask, TokenPoolingType.all).tolist() == mask.tolist()


def test_event_bus_and_reflection():
    from d9d_b200.loop.event import Event, EventBus, subscribe, subscribe_annotated

    ev_a, ev_b = Event[int]("a"), Event[int]("b")
    seen = []
    bus = EventBus()
    bus.subscribe(ev_a, lambda x: seen.append(("fn", x)))

    class Listener:
        @subscribe(ev_a)
        def on_a(self, x):
            seen.append(("a", x))

        @subscribe(ev_b)
        def on_b(self, x):
            seen.append(("b", x))

    subscribe_annotated(bus, Listener())
    bus.trigger(ev_a, 1)
    with bus.bounded(ev_a, ev_b, 2):
        seen.append("body")
    assert seen == [("fn", 1), ("a", 1), ("fn", 2), ("a", 2), "body", ("b", 2)]
    with pytest.raises(RuntimeError):
        with bus.bounded(ev_a, ev_b, 3):
            raise RuntimeError("boom")
    assert seen[-1] == ("a", 3)  # post event is not fired on failure


def test_stepper_periodic_actions():
    from d9d_b200.loop.component import Stepper

    s = Stepper(initial_step=0, total_steps=10)
    fired = []
    while s.current_step < s.total_steps:
        if s.should_do_action(3, enable_on_last_step_if_periodic=True):
            fired.append(s.current_step)
        s.step()
    assert fired == [2, 5, 8, 9]
    assert not Stepper(0, 10).should_do_action("disable") and Stepper(9, 10).should_do_action("last_step")
    restored = Stepper(0, 10)
    restored.load_state_dict(s.state_dict())
    assert restored.current_step == 10
