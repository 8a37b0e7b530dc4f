"""Metrics, trackers, PEFT, model-state IO, pipeline state, the inference loop."""

import json
import re

import pytest
import torch
from torch import nn


# ------------------------------------------------------------------------------------------------ metrics
def test_aggregation_and_compose_metrics():
    from d9d_b200.metric.impl.aggregation import SumMetric, WeightedMeanMetric
    from d9d_b200.metric.impl.container import ComposeMetric

    m = ComposeMetric({"tokens": SumMetric(), "loss": WeightedMeanMetric()})
    m.children["tokens"].update(torch.tensor([3.0, 4.0]))
    m.children["loss"].update(torch.tensor([1.0, 3.0]), torch.tensor([1.0, 3.0]))
    m.children["loss"].update(torch.tensor(5.0), torch.tensor(0.0))
    out = m.compute()
    assert float(out["tokens"]) == 7.0 and abs(float(out["loss"]) - 2.5) < 1e-6
    state = m.state_dict()
    m2 = ComposeMetric({"tokens": SumMetric(), "loss": WeightedMeanMetric()})
    m2.load_state_dict(state)
    assert float(m2.compute()["tokens"]) == 7.0
    m.reset()
    assert float(m.compute()["tokens"]) == 0.0


def test_classification_metrics_match_sklearn():
    sk = pytest.importorskip("sklearn.metrics")
    from d9d_b200.metric.impl.classification import BinaryAUROCMetric, confusion_matrix_metric

    g = torch.Generator().manual_seed(0)
    logits = torch.randn(500, 5, generator=g)
    labels = torch.randint(0, 5, (500,), generator=g)
    preds = logits.argmax(-1)
    for avg, builder in (("macro", lambda b: b.macro()), ("micro", lambda b: b.micro()), ("weighted", lambda b: b.weighted())):
        metric = builder(confusion_matrix_metric().multiclass(num_classes=5).with_f1()).build()
        for chunk in range(0, 500, 100):  # streaming updates
            metric.update(logits[chunk:chunk + 100], labels[chunk:chunk + 100])
        want = sk.f1_score(labels.numpy(), preds.numpy(), average=avg)
        assert abs(float(metric.compute()) - want) < 1e-5, avg
    per_class = confusion_matrix_metric().multiclass(num_classes=5).with_recall().per_class().build()
    per_class.update(logits, labels)
    torch.testing.assert_close(per_class.compute().double(), torch.tensor(sk.recall_score(labels.numpy(), preds.numpy(), average=None)),
                               rtol=1e-5, atol=1e-6)
    acc = confusion_matrix_metric().binary(threshold=0.5).with_accuracy().build()
    probs = torch.rand(300, generator=g)
    y = (torch.rand(300, generator=g) < probs).long()
    acc.update(probs, y)
    assert abs(float(acc.compute()) - sk.accuracy_score(y.numpy(), (probs > 0.5).long().numpy())) < 1e-6
    auroc = BinaryAUROCMetric(num_bins=20000)
    auroc.update(probs, y)
    assert abs(float(auroc.compute()) - sk.roc_auc_score(y.numpy(), probs.numpy())) < 2e-3
    with pytest.raises(ValueError):
        confusion_matrix_metric().binary().multiclass(3)


# ------------------------------------------------------------------------------------------------ trackers
def test_jsonl_tracker_resumes_into_the_same_file(tmp_path):
    from d9d_b200.tracker import RunConfig, tracker_from_config
    from d9d_b200.tracker.provider.jsonl import JsonlTrackerConfig
    from d9d_b200.tracker.provider.null import NullTrackerConfig

    cfg = JsonlTrackerConfig(directory=str(tmp_path))
    tracker = tracker_from_config(cfg)
    with tracker.open(RunConfig(name="r", description=None, hparams={"lr": 1e-3})) as run:
        run.set_context({"stage": "train"})
        run.set_step(3)
        run.scalar("loss", 1.5)
        run.bins("hist", torch.arange(4.0), context={"layer": "0"})
    state = tracker.state_dict()
    resumed = tracker_from_config(cfg)
    resumed.load_state_dict(state)
    with resumed.open(RunConfig(name="r", description=None)) as run:
        run.set_step(4)
        run.scalar("loss", 1.2)
    files = list(tmp_path.glob("*.jsonl"))
    assert len(files) == 1
    recs = [json.loads(line) for line in files[0].read_text().splitlines()]
    losses = [(r["step"], r["value"]) for r in recs if r.get("name") == "loss"]
    assert losses == [(3, 1.5), (4, 1.2)]
    assert any(r.get("name") == "hist" and r.get("context", {}).get("layer") == "0" for r in recs)
    with tracker_from_config(NullTrackerConfig()).open(RunConfig(name="x", description=None)) as run:
        run.scalar("anything", 1.0)  # accepted and dropped


# ------------------------------------------------------------------------------------------------ PEFT
class _Block(nn.Module):
    def __init__(self):
        super().__init__()
        from d9d_b200.module.block.moe import GroupedLinear

        self.proj = nn.Linear(8, 8, bias=False)
        self.other = nn.Linear(8, 4, bias=False)
        self.experts = GroupedLinear(3, 8, 6)

    def forward(self, x):
        return self.other(self.proj(x))


def test_lora_inject_train_merge():
    from d9d_b200.peft import inject_peft_and_freeze, merge_peft
    from d9d_b200.peft.all import peft_method_from_config
    from d9d_b200.peft.all.config import PeftStackConfig
    from d9d_b200.peft.lora import LoRAGroupedLinear, LoRALinear

    torch.manual_seed(0)
    model = _Block()
    x = torch.randn(5, 8)
    before = model(x).detach()
    cfg = PeftStackConfig.model_validate({"kind": "stack", "methods": [
        {"kind": "lora", "module_name_pattern": r"proj|experts", "params": {"r": 2, "alpha": 4, "dropout": 0.0}},
        {"kind": "full_tune", "module_name_pattern": r"other"}]})
    method = peft_method_from_config(cfg)
    mapper = inject_peft_and_freeze(method, model)
    assert isinstance(model.proj, LoRALinear) and isinstance(model.experts, LoRAGroupedLinear)
    trainable = {n for n, p in model.named_parameters() if p.requires_grad}
    assert trainable == {"proj.lora_A.weight", "proj.lora_B.weight", "experts.lora_A.weight", "experts.lora_B.weight", "other.weight"}, trainable
    renames = {(tuple(g.inputs)[0], tuple(g.outputs)[0]) for g in mapper.state_dependency_groups()}
    assert ("proj.weight", "proj.base.weight") in renames and ("experts.weight", "experts.base.weight") in renames
    torch.testing.assert_close(model(x), before)  # B starts at zero: injection does not change the function
    with torch.no_grad():
        model.proj.lora_B.weight.normal_()
    changed = model(x).detach()
    assert not torch.allclose(changed, before)
    merge_peft(method, model)
    assert isinstance(model.proj, nn.Linear) and not isinstance(model.proj, LoRALinear)
    torch.testing.assert_close(model(x), changed, rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------------------------------------ model state IO
def test_save_and_load_model_state_through_mappers(tmp_path):
    from d9d_b200.model_state.io import load_model_state, read_model_state, save_model_state
    from d9d_b200.model_state.mapper.adapters import identity_mapper_from_module
    from d9d_b200.model_state.mapper.compose import ModelStateMapperParallel, ModelStateMapperSequential
    from d9d_b200.model_state.mapper.leaf import (ModelStateMapperConcatenateTensors, ModelStateMapperIdentity,
                                                  ModelStateMapperRename, ModelStateMapperTranspose)

    torch.manual_seed(0)
    src = nn.Sequential(nn.Linear(4, 6), nn.Linear(6, 3))
    save_model_state(tmp_path / "plain", identity_mapper_from_module(src), src, shard_size_gb=1e-7, show_progress=False)
    index = json.loads((tmp_path / "plain" / "model.safetensors.index.json").read_text())
    assert set(index["weight_map"]) == set(src.state_dict()) and len(set(index["weight_map"].values())) > 1  # really sharded

    dst = nn.Sequential(nn.Linear(4, 6), nn.Linear(6, 3))
    load_model_state(tmp_path / "plain", identity_mapper_from_module(dst), "cpu", dst, show_progress=False)
    for k, v in src.state_dict().items():
        torch.testing.assert_close(dst.state_dict()[k], v)

    # read through a mapper DAG: fuse both biases, transpose + rename a weight, drop everything else
    mapper = ModelStateMapperParallel([
        ModelStateMapperConcatenateTensors(["0.bias", "1.bias"], "all_bias", dim=0),
        ModelStateMapperSequential([ModelStateMapperTranspose("0.weight", dims=(0, 1)), ModelStateMapperRename("0.weight", "w0_t")]),
        ModelStateMapperIdentity("1.weight"),
    ])
    got = dict(read_model_state(tmp_path / "plain", mapper, "cpu", show_progress=False))
    assert set(got) == {"all_bias", "w0_t", "1.weight"}
    torch.testing.assert_close(got["all_bias"], torch.cat([src[0].bias, src[1].bias]))
    torch.testing.assert_close(got["w0_t"], src[0].weight.t())
    with pytest.raises(ValueError):
        list(read_model_state(tmp_path / "plain", ModelStateMapperIdentity("missing.key"), "cpu", show_progress=False))


# ------------------------------------------------------------------------------------------------ pipeline state
def test_pipeline_state_global_and_shard_views():
    from d9d_b200.internals.pipeline_state import PipelineStateHandler

    h = PipelineStateHandler(sharding_spec={}, num_shards=2)
    g = h.global_state()
    g["labels"] = torch.arange(8).view(4, 2)
    g["names"] = ["a", "b", "c", "d"]
    assert h.sharded_state(1)["labels"].tolist() == [[4, 5], [6, 7]] and h.sharded_state(0)["names"] == ["a", "b"]
    for mb in range(2):
        h.sharded_state(mb)["loss"] = torch.tensor(float(mb + 1))  # scalars written per microbatch are stacked
        h.sharded_state(mb)["hidden"] = torch.full((2, 3), float(mb))
    assert g["loss"].tolist() == [1.0, 2.0] and g["hidden"].shape == (4, 3)
    assert "loss" in g and "nope" not in g
    h.reset()
    assert "loss" not in h.global_state()


# ------------------------------------------------------------------------------------------------ inference loop
def test_inference_loop_runs_the_task_over_every_batch(tmp_path):
    from d9d_b200.core.dist_context import DeviceMeshParameters
    from d9d_b200.loop.config import InferenceConfig
    from d9d_b200.loop.control import BuildForwardInputsResult, InferenceTask
    from d9d_b200.loop.run import InferenceConfigurator
    from tests.helpers_train import LMProvider, SyntheticDataProvider, dense_params

    class Collect(InferenceTask):
        def __init__(self):
            self.batches = 0
            self.tokens = 0

        def build_forward_inputs(self, ctx):
            ctx.state["labels"] = ctx.batch["labels"]
            return BuildForwardInputsResult(inputs={"input_ids": ctx.batch["input_ids"]},
                                            kwargs={"labels": ctx.batch["labels"], "position_ids": ctx.batch["position_ids"]})

        def process_outputs(self, ctx):
            assert ctx.pipeline_results["logps"].shape == ctx.state["labels"].shape
            self.batches += 1
            self.tokens += ctx.state["labels"].numel()

    task = Collect()
    cfg = InferenceConfig.model_validate({
        "batching": {"global_batch_size": 4, "microbatch_size": 4},
        "data_loading": {"num_workers": 0, "pin_memory": False, "persistent_workers": False},
        "model_stage_factory": {"source_checkpoint": None, "checkpoint_only_trainable_parameters": False},
        "determinism": {"base_seed": 1}, "gc": {"period_steps": 100},
        "checkpointing": {"save_dir": str(tmp_path / "ckpt"), "period_steps": "disable", "num_to_keep": 1},
        "profiling": None})
    job = InferenceConfigurator(mesh=DeviceMeshParameters(), parameters=cfg, task_provider=lambda ctx: task,
                                model_provider=LMProvider(dense_params()), data_provider=SyntheticDataProvider(num_samples=18)).configure()
    job.infer()
    assert task.batches == 5 and task.tokens == 18 * 16  # the ragged last batch is kept


def test_aim_tracker_against_a_stub_backend(monkeypatch):
    """The Aim adapter: run properties, step / context stamping, histogram conversion, run-hash resume
    (reference ``test/d9d_test/tracker/test_aim.py``) - exercised against a stand-in ``aim`` module."""
    import sys
    import types

    from d9d_b200.tracker import RunConfig
    from d9d_b200.tracker.factory import tracker_from_config
    from d9d_b200.tracker.provider.aim.config import AimConfig

    config = AimConfig(repo="/tmp/aim-repo", log_system_params=False, capture_terminal_logs=False)
    monkeypatch.delitem(sys.modules, "aim", raising=False)
    import importlib.util

    if importlib.util.find_spec("aim") is None:
        with pytest.raises(ImportError):
            tracker_from_config(config)

    opened = []

    class FakeRun(dict):
        def __init__(self, run_hash=None, **kwargs):
            super().__init__()
            self.hash = run_hash or f"hash-{len(opened)}"
            self.kwargs, self.tracked, self.closed = kwargs, [], False
            opened.append(self)

        def track(self, value, name, step, context):
            self.tracked.append((name, value, step, dict(context)))

        def close(self):
            self.closed = True

    class FakeDistribution:
        def __init__(self, hist, bin_range):
            self.hist, self.bin_range = hist, bin_range

    monkeypatch.setitem(sys.modules, "aim", types.SimpleNamespace(Run=FakeRun, Distribution=FakeDistribution))
    tracker = tracker_from_config(config)
    assert tracker.state_dict() == {"restart_hash": None}
    with tracker.open(RunConfig(name="exp", description="d", hparams={"lr": 1e-3})) as run:
        run.set_context({"stage": "train"})
        run.set_step(7)
        run.scalar("loss", 0.5)
        run.scalar("loss", 0.25, context={"subset": "val"})
        run.bins("tokens_per_expert", torch.tensor([3, 0, 5]))
    backend = opened[0]
    assert backend.closed and backend.name == "exp" and backend["hparams"] == {"lr": 1e-3}
    assert backend.kwargs["repo"] == "/tmp/aim-repo" and backend.kwargs["log_system_params"] is False
    assert backend.tracked[0] == ("loss", 0.5, 7, {"stage": "train"})
    assert backend.tracked[1] == ("loss", 0.25, 7, {"stage": "train", "subset": "val"})
    name, dist, step, ctx = backend.tracked[2]
    assert name == "tokens_per_expert" and list(dist.hist) == [3, 0, 5] and dist.bin_range == (0, 3) and step == 7

    state = tracker.state_dict()
    assert state == {"restart_hash": "hash-0"}
    resumed = tracker_from_config(config)
    resumed.load_state_dict(state)
    with resumed.open(RunConfig(name="exp", description=None)):
        pass
    assert opened[1].hash == "hash-0"  # the restarted job appends to the same run


def test_throughput_meter_reports_tokens_per_second_and_mfu():
    import time

    from d9d_b200.loop.component import Stepper
    from d9d_b200.loop.event import EventBus
    from d9d_b200.loop.event.catalogue.train import (EVENT_TRAIN_READY, EVENT_TRAIN_STEP_POST, EVENT_TRAIN_STEP_PRE, EventStepContext,
                                                     EventTrainReadyContext)
    from d9d_b200.recipes import ThroughputMeter, transformer_flops_per_token

    class Run:
        def __init__(self):
            self.logged = []

        def scalar(self, name, value, context=None):
            self.logged.append((name, value))

    flops = transformer_flops_per_token(num_parameters_active=1_000_000, num_layers=2, hidden_size=64, seq_len=128)
    assert flops == 6e6 + 12 * 2 * 64 * 128 / 2
    bus, run = EventBus(), Run()
    meter = ThroughputMeter(tokens_per_step=1000, world_size=2, flops_per_token=flops, peak_flops_per_device=1e12, period_steps=3,
                            skip_first_steps=1)
    meter.install(bus)
    bus.trigger(EVENT_TRAIN_READY, EventTrainReadyContext(run=run))
    step = EventStepContext(stepper=Stepper(0, 10))
    for _ in range(6):
        bus.trigger(EVENT_TRAIN_STEP_PRE, step)
        time.sleep(0.01)
        bus.trigger(EVENT_TRAIN_STEP_POST, step)
    meter.flush()
    assert len(meter.history) == 2  # steps 2-4 after the skipped first one, then the rest on flush
    for entry in meter.history:
        assert 0.009 < entry["step_seconds"] < 0.05
        assert abs(entry["tokens_per_second"] - 1000 / entry["step_seconds"]) < 1e-6
        assert abs(entry["mfu"] - 1000 * flops / entry["step_seconds"] / 2e12) < 1e-9
    assert {name for name, _ in run.logged} == {"throughput/tokens_per_second", "throughput/step_seconds", "throughput/mfu"}


def test_lora_layers_initialise_their_frozen_base_after_meta_construction():
    """A stage is built on the meta device, adapters are injected there, and only then is memory allocated: the late
    initialisation has to cover the frozen base weights too (they stay garbage otherwise when no checkpoint is loaded)."""
    from d9d_b200.module.block.moe import GroupedLinear
    from d9d_b200.peft import inject_peft_and_freeze
    from d9d_b200.peft.all import peft_method_from_config
    from d9d_b200.peft.lora.config import LoRAConfig

    class Tiny(nn.Module):
        def __init__(self):
            super().__init__()
            self.proj = nn.Linear(8, 8, bias=False)
            self.experts = GroupedLinear(2, 8, 8)

        def reset_parameters(self):
            self.proj.reset_parameters()
            self.experts.reset_parameters()

    with torch.device("meta"):
        model = Tiny()
        inject_peft_and_freeze(peft_method_from_config(LoRAConfig.model_validate(
            {"kind": "lora", "module_name_pattern": "proj|experts", "params": {"r": 2, "alpha": 4, "dropout": 0.0}})), model)
    model.to_empty(device="cpu")
    with torch.no_grad():
        for p in model.parameters():
            p.fill_(float("nan"))  # what uninitialised memory may look like
        model.reset_parameters()
    assert all(torch.isfinite(p).all() for p in model.parameters())
    assert float(model.proj.lora_B.weight.abs().sum()) == 0.0 and float(model.proj.base.weight.abs().sum()) > 0.0
