"""End-to-end Trainer: trains on CPU, loss goes down, checkpoint/resume is exact, export round-trips; gloo DP+PP."""

import json

import pytest
import torch

from tests.dist_utils import run_distributed
from tests.helpers_train import LMProvider, SFTTask, SyntheticDataProvider, dense_params, moe_params, trainer_config


def _make_trainer(tmp, mesh=None, moe=False, schedule=None, ckpt_period="disable", total_batch=8, micro=4, log=True, samples=64,
                  optimizer=None, fold_scaling=True, dtype=torch.float32, source=None, async_save=False, recompute=False, layers=2):
    from d9d_b200.core.dist_context import DeviceMeshParameters
    from d9d_b200.loop.auto import AutoLRSchedulerProvider, AutoOptimizerProvider
    from d9d_b200.loop.auto.auto_lr_scheduler import PiecewiseConfig
    from d9d_b200.loop.auto.auto_optimizer import AdamWOptimizerConfig
    from d9d_b200.loop.run import TrainingConfigurator

    sched = PiecewiseConfig.model_validate({"name": "piecewise", "scheduler": {"initial_multiplier": 0.1, "phases": [
        {"mode": "percentage", "percentage": 0.25, "target_multiplier": 1.0, "curve": {"type": "linear"}},
        {"mode": "rest", "target_multiplier": 0.1, "curve": {"type": "cosine"}}]}})
    return TrainingConfigurator(
        mesh=mesh or DeviceMeshParameters(),
        parameters=trainer_config(tmp, total_batch=total_batch, micro=micro, schedule=schedule, ckpt_period=ckpt_period,
                                  log_dir=(tmp / "logs") if log else None, source=source, async_save=async_save).model_copy(update={}, deep=True)
        if fold_scaling else _without_folding(trainer_config(tmp, total_batch=total_batch, micro=micro, schedule=schedule,
                                                             ckpt_period=ckpt_period, log_dir=(tmp / "logs") if log else None,
                                                             source=source)),
        task_provider=lambda ctx: SFTTask(ctx.dist_context),
        model_provider=LMProvider(moe_params(layers) if moe else dense_params(layers), moe=moe, dtype=dtype,
                                  activation_checkpointing=recompute),
        data_provider=SyntheticDataProvider(num_samples=samples),
        optimizer_provider=AutoOptimizerProvider(optimizer or AdamWOptimizerConfig(lr=3e-3, weight_decay=0.0)),
        lr_scheduler_provider=AutoLRSchedulerProvider(sched),
    ).configure()


def _without_folding(cfg):
    cfg = cfg.model_copy(deep=True)
    cfg.gradient_manager.fold_scaling_into_optimizer = False
    return cfg


def _read_losses(tmp):
    files = list((tmp / "logs").glob("*.jsonl"))
    assert len(files) == 1, files
    recs = [json.loads(line) for line in files[0].read_text().splitlines()]
    return {r["step"]: r["value"] for r in recs if r.get("name") == "loss"}, recs


@pytest.mark.parametrize("moe", [False, True])
def test_train_local_loss_decreases(tmp_path, moe):
    trainer = _make_trainer(tmp_path, moe=moe)
    assert trainer.state.stepper.total_steps == 8  # 64 samples / global batch 8
    trainer.train()
    losses, recs = _read_losses(tmp_path)
    steps = sorted(losses)
    assert steps == list(range(trainer.state.stepper.total_steps))
    assert losses[steps[-1]] < losses[steps[0]]
    assert any(r.get("name") == "l2_grad_norm_total" for r in recs)
    assert any(r.get("name") == "num_tokens" for r in recs)


def test_resume_is_exact(tmp_path):
    # reference run: all steps in one go
    full = _make_trainer(tmp_path / "a", ckpt_period="disable")
    full.train()
    ref_state = {k: v.clone() for k, v in full.state.tracked_modules.modules[0].state_dict().items()}

    # interrupted run: stop after 3 steps (checkpoint every step), then resume with a fresh Trainer
    class _Stop(Exception):
        pass

    part = _make_trainer(tmp_path / "b", ckpt_period=1)
    from d9d_b200.loop.event.catalogue.train import EVENT_TRAIN_STEP_PRE

    def stop_at_3(ctx):
        if ctx.stepper.current_step == 3:
            raise _Stop

    part.state.event_bus.subscribe(EVENT_TRAIN_STEP_PRE, stop_at_3)
    with pytest.raises(_Stop):
        part.train()
    assert sorted(p.name for p in (tmp_path / "b" / "ckpt" / "t").iterdir()) == ["save-2", "save-3"]  # keep-last-2 rotation

    resumed = _make_trainer(tmp_path / "b", ckpt_period=1)
    resumed.train()
    assert resumed.state.stepper.current_step == resumed.state.stepper.total_steps
    for k, v in resumed.state.tracked_modules.modules[0].state_dict().items():
        torch.testing.assert_close(v, ref_state[k], rtol=0, atol=0, msg=lambda m, k=k: f"{k}: {m}")

    # export -> reload round trip
    from d9d_b200.model_state.io import load_model_state
    from d9d_b200.model_state.mapper.adapters import identity_mapper_from_module

    resumed.export(tmp_path / "export", load_checkpoint=False)
    assert (tmp_path / "export" / "model.safetensors.index.json").exists()
    fresh = _make_trainer(tmp_path / "c").state.tracked_modules.modules[0]
    load_model_state(tmp_path / "export", identity_mapper_from_module(fresh), "cpu", fresh, show_progress=False)
    for k, v in fresh.state_dict().items():
        torch.testing.assert_close(v, ref_state[k], rtol=0, atol=0)


def _dist_worker(rank, world, tmp, mesh_kwargs, schedule, moe, source=None, recompute=False, layers=2, samples=32):
    from pathlib import Path

    from d9d_b200.core.dist_context import DeviceMeshParameters

    tmp = Path(tmp)
    trainer = _make_trainer(tmp, mesh=DeviceMeshParameters(**mesh_kwargs), moe=moe, schedule=schedule, total_batch=8, micro=2,
                            log=True, samples=samples, source=source, recompute=recompute, layers=layers)
    trainer.train()
    trainer.export(tmp / "export", load_checkpoint=False)
    if rank == 0:
        losses, _ = _read_losses(tmp)
        steps = sorted(losses)
        assert losses[steps[-1]] < losses[steps[0]], losses


@pytest.mark.dist
@pytest.mark.parametrize("mesh_kwargs,schedule,moe,world", [
    ({"data_parallel_replicate": 2}, {"schedule": "gpipe"}, False, 2),
    ({"pipeline_parallel": 2}, {"schedule": "1f1b", "num_stages_per_rank": 1, "zero_bubble": True}, False, 2),
    ({"pipeline_parallel": 2, "data_parallel_replicate": 2}, {"schedule": "looped_bfs", "num_stages_per_rank": 1}, True, 4),
])
def test_train_distributed_gloo(tmp_path, mesh_kwargs, schedule, moe, world):
    run_distributed(_dist_worker, world, str(tmp_path), mesh_kwargs, schedule, moe)
    assert (tmp_path / "export" / "model.safetensors.index.json").exists()


def test_dp_matches_single_process(tmp_path):
    """2-way replicate training must reproduce the single-process run (same global batch) step for step."""
    single = _make_trainer(tmp_path / "s", total_batch=8, micro=2, samples=32)
    single.train()
    ref, _ = _read_losses(tmp_path / "s")
    run_distributed(_dist_worker, 2, str(tmp_path / "d"), {"data_parallel_replicate": 2}, {"schedule": "gpipe"}, False)
    got, _ = _read_losses(tmp_path / "d")
    assert sorted(ref) == sorted(got)
    # samples are distributed differently (round-robin sharding) so per-step batches differ; the totals must be close
    # and the first step (same init, different but same-sized batch) must be within noise
    assert abs(ref[0] - got[0]) < 0.2


def _async_worker(rank, world, tmp):
    from pathlib import Path

    from d9d_b200.core.dist_context import DeviceMeshParameters

    trainer = _make_trainer(Path(tmp), mesh=DeviceMeshParameters(data_parallel_replicate=2), ckpt_period=2, total_batch=8, micro=2,
                            samples=32, async_save=True)
    trainer.train()


def test_asynchronous_checkpoints_and_interrupted_saves(tmp_path):
    """``async_save``: writes overlap training and are complete (and rotated) when ``train()`` returns; a save directory
    without DCP's ``.metadata`` (a job killed mid-save) is ignored on resume."""
    import shutil

    trainer = _make_trainer(tmp_path / "a", ckpt_period=3, async_save=True)
    trainer.train()
    saved = tmp_path / "a" / "ckpt" / "t"
    assert sorted(p.name for p in saved.iterdir()) == ["save-6", "save-8"]  # 8 steps, period 3 + last step, keep 2
    assert all((p / ".metadata").exists() for p in saved.iterdir())

    # a later, half-written checkpoint must not be picked up; the job is complete at save-8 and does nothing
    shutil.copytree(saved / "save-8", saved / "save-9")
    (saved / "save-9" / ".metadata").unlink()
    again = _make_trainer(tmp_path / "a", ckpt_period=3, async_save=True)
    again.train()
    assert again.state.stepper.current_step == 8

    # resume after an interruption reproduces the uninterrupted weights also with background saves
    reference = {k: v.clone() for k, v in trainer.state.tracked_modules.modules[0].state_dict().items()}
    (saved / "save-8" / ".metadata").unlink()  # pretend the last save never finished -> resume from save-6
    resumed = _make_trainer(tmp_path / "a", ckpt_period=3, async_save=True)
    resumed.train()
    for k, v in resumed.state.tracked_modules.modules[0].state_dict().items():
        torch.testing.assert_close(v, reference[k], rtol=0, atol=0, msg=lambda m, k=k: f"{k}: {m}")

    run_distributed(_async_worker, 2, str(tmp_path / "dist"))  # background saves with a process group
    assert sorted(p.name for p in (tmp_path / "dist" / "ckpt" / "t").iterdir()) == ["save-2", "save-4"]


def _jobs_worker(rank, world, jobs):
    """Several training jobs over one set of processes (spawning processes dominates the cost of these tests)."""
    for job in jobs:
        _dist_worker(rank, world, *job)


def _assert_same_trajectory(reference_dir, job_dir, label):
    ref, _ = _read_losses(reference_dir)
    got, _ = _read_losses(job_dir)
    assert sorted(ref) == sorted(got), label
    for step in ref:
        assert abs(ref[step] - got[step]) < 2e-3 * max(1.0, abs(ref[step])), (label, step, ref[step], got[step])


_SEQUENCE_SHARDED = {
    2: [{"context_parallel_shard": 2}, {"tensor_parallel": 2}],
    4: [{"context_parallel_replicate": 2, "pipeline_parallel": 2},
        {"tensor_parallel": 2, "context_parallel_replicate": 2},
        {"tensor_parallel": 2, "context_parallel_shard": 2},
        # FSDP-sharded stages under a zero-bubble schedule (whole backward in the I slot)
        {"pipeline_parallel": 2, "context_parallel_shard": 2, "zero_bubble": True}],
}


@pytest.mark.dist
@pytest.mark.parametrize("world", [2, 4])
def test_sequence_sharded_training_reproduces_the_single_process_run(tmp_path, world):
    """Ranks of a context- / tensor-parallel group read the same samples and split every sequence (and, for tensor
    parallelism, the heads and MLP columns): the loss trajectory must equal the single-process one (same batches, same
    maths - only the reduction order differs)."""
    _make_trainer(tmp_path / "init", log=False).export(tmp_path / "weights", load_checkpoint=False)  # shared initial weights
    single = _make_trainer(tmp_path / "s", total_batch=8, micro=2, samples=32, source=tmp_path / "weights")
    single.train()
    jobs = []
    for i, mesh_kwargs in enumerate(_SEQUENCE_SHARDED[world]):
        mesh_kwargs = dict(mesh_kwargs)
        zero_bubble = mesh_kwargs.pop("zero_bubble", False)
        schedule = ({"schedule": "1f1b", "num_stages_per_rank": 1, "zero_bubble": zero_bubble} if "pipeline_parallel" in mesh_kwargs
                    else {"schedule": "gpipe"})
        jobs.append((str(tmp_path / f"job{i}"), mesh_kwargs, schedule, False, str(tmp_path / "weights")))
    run_distributed(_jobs_worker, world, jobs)
    for i, mesh_kwargs in enumerate(_SEQUENCE_SHARDED[world]):
        _assert_same_trajectory(tmp_path / "s", tmp_path / f"job{i}", mesh_kwargs)


def test_resume_is_exact_with_stochastic_adamw(tmp_path):
    """The stochastic-rounding optimizer through a DCP checkpoint: moments, the *integer* step counters (which DCP hands back
    under stringified parameter ids) and the rounding generator must all resume, otherwise the trajectory drifts."""
    from d9d_b200.loop.auto.auto_optimizer import StochasticAdamWOptimizerConfig
    from d9d_b200.loop.event.catalogue.train import EVENT_TRAIN_STEP_PRE

    def make(tmp):
        return _make_trainer(tmp, ckpt_period=2, dtype=torch.bfloat16,
                             optimizer=StochasticAdamWOptimizerConfig(lr=3e-3, weight_decay=0.01, state_dtype="float32"))

    full = make(tmp_path / "full")
    full.train()

    class _Stop(Exception):
        pass

    def stop(ctx):
        if ctx.stepper.current_step == 5:
            raise _Stop

    part = make(tmp_path / "cut")
    part.state.event_bus.subscribe(EVENT_TRAIN_STEP_PRE, stop)
    with pytest.raises(_Stop):
        part.train()
    resumed = make(tmp_path / "cut")
    resumed.train()
    optimizer = resumed.state.optimizer.optimizers[0]
    assert {int(s["step"]) for s in optimizer.state.values()} == {8}
    want = full.state.tracked_modules.modules[0].state_dict()
    for k, v in resumed.state.tracked_modules.modules[0].state_dict().items():
        assert torch.equal(v, want[k]), k


def test_folding_the_gradient_scale_into_the_optimizer_changes_nothing(tmp_path):
    """1/sum(w) and the clip coefficient handed to StochasticAdamW as a device scalar == rewriting the gradients."""
    from d9d_b200.loop.auto.auto_optimizer import StochasticAdamWOptimizerConfig

    finals = []
    for fold in (True, False):
        opt = StochasticAdamWOptimizerConfig(lr=3e-3, weight_decay=0.0, state_dtype="float32")
        trainer = _make_trainer(tmp_path / f"fold{fold}", optimizer=opt, fold_scaling=fold, log=False, dtype=torch.bfloat16)
        assert (trainer.state.gradient_manager.pending_scale is not None) == fold
        trainer.train()
        finals.append({k: v.detach().clone() for k, v in trainer.state.tracked_modules.modules[0].state_dict().items()})
    # identical maths up to fp32 association order; a handful of stochastic-rounding decisions may flip by one bf16 ulp
    a = torch.cat([v.float().flatten() for v in finals[0].values()])
    b = torch.cat([finals[1][k].float().flatten() for k in finals[0]])
    assert float((a != b).float().mean()) < 0.02
    torch.testing.assert_close(a, b, rtol=2e-2, atol=2e-2)


def _resume_worker(rank, world, tmp, mesh_kwargs, schedule, moe, stop_at):
    """Train with a checkpoint every 2 steps; ``stop_at``: simulate a crash right before that step."""
    from pathlib import Path

    from d9d_b200.core.dist_context import DeviceMeshParameters
    from d9d_b200.loop.event.catalogue.train import EVENT_TRAIN_STEP_PRE

    class _Crash(Exception):
        pass

    tmp = Path(tmp)
    trainer = _make_trainer(tmp, mesh=DeviceMeshParameters(**mesh_kwargs), moe=moe, schedule=schedule, total_batch=8, micro=2, log=True,
                            samples=64, ckpt_period=2, source=tmp.parent / "weights")
    if stop_at is not None:
        def crash(ctx):
            if ctx.stepper.current_step == stop_at:
                torch.distributed.barrier()  # let every rank finish the collectives of the previous step before anyone leaves
                raise _Crash

        trainer.state.event_bus.subscribe(EVENT_TRAIN_STEP_PRE, crash)
        try:
            trainer.train()
        except _Crash:
            return
        raise AssertionError("the simulated crash did not happen")
    trainer.train()
    trainer.export(tmp / "export", load_checkpoint=False)


_RESUME_CASES = [
    ("pp_fsdp", {"pipeline_parallel": 2, "data_parallel_shard": 2}, {"schedule": "1f1b", "num_stages_per_rank": 1, "zero_bubble": True}, False),
    ("hsdp_ep", {"data_parallel_replicate": 2, "data_parallel_shard": 2, "expert_parallel": 2}, {"schedule": "gpipe"}, True),
]


def _resume_phase(rank, world, root, phase, stop_at):
    """One phase (uninterrupted run / run until the crash / restart) of every case over one set of processes."""
    from pathlib import Path

    for name, mesh_kwargs, schedule, moe in _RESUME_CASES:
        _resume_worker(rank, world, str(Path(root) / name / phase), mesh_kwargs, schedule, moe, stop_at)


@pytest.mark.dist
def test_distributed_resume_is_exact(tmp_path):
    """Kill a 4-rank job (pipeline x FSDP with a zero-bubble schedule; HSDP x expert parallel MoE) after step 5, restart it in
    fresh processes: sharded DCP checkpoints, per-rank data-loader positions and per-stage optimizer states must bring it
    back onto the exact trajectory of an uninterrupted job."""
    from d9d_b200.model_state.io import read_model_state
    from d9d_b200.model_state.mapper.adapters import identity_mapper_from_module

    inits = {}
    for name, _, _, moe in _RESUME_CASES:
        inits[name] = _make_trainer(tmp_path / name / "init", moe=moe, log=False)
        inits[name].export(tmp_path / name / "weights", load_checkpoint=False)
    run_distributed(_resume_phase, 4, str(tmp_path), "full", None)
    run_distributed(_resume_phase, 4, str(tmp_path), "cut", 5)
    for name, *_ in _RESUME_CASES:
        assert sorted(p.name for p in (tmp_path / name / "cut" / "ckpt" / "t").iterdir()) == ["save-2", "save-4"]
    run_distributed(_resume_phase, 4, str(tmp_path), "cut", None)  # fresh processes: the restart

    for name, *_ in _RESUME_CASES:
        full_losses, _ = _read_losses(tmp_path / name / "full")
        cut_losses, _ = _read_losses(tmp_path / name / "cut")
        for step in range(4, 8):  # steps replayed after the restart
            assert cut_losses[step] == full_losses[step], (name, step, cut_losses[step], full_losses[step])
        mapper = identity_mapper_from_module(inits[name].state.tracked_modules.modules[0])
        want = dict(read_model_state(tmp_path / name / "full" / "export", mapper, "cpu", show_progress=False))
        got = dict(read_model_state(tmp_path / name / "cut" / "export", mapper, "cpu", show_progress=False))
        assert want.keys() == got.keys()
        for key in want:
            torch.testing.assert_close(got[key], want[key], rtol=0, atol=0, msg=lambda m, key=key: f"{name} {key}: {m}")  # noqa: B023


_PIPELINE_COMBINATIONS = [
    # expert-parallel all-to-alls inside the split backward of a zero-bubble schedule, activation recomputation
    ({"pipeline_parallel": 2, "context_parallel_replicate": 2, "expert_parallel": 2},
     {"schedule": "1f1b", "num_stages_per_rank": 1, "zero_bubble": True}, True),
    # interleaved (two stages per rank) schedule over FSDP-sharded stages, activation recomputation
    ({"pipeline_parallel": 2, "context_parallel_shard": 2}, {"schedule": "looped_bfs", "num_stages_per_rank": 2}, False),
    ({"pipeline_parallel": 2, "context_parallel_replicate": 2, "expert_parallel": 2}, {"schedule": "dual_pipe_v"}, True),
    # split backward through real decoder layers: interleaved zero-bubble 1F1B and the V-shaped zero-bubble schedule
    ({"pipeline_parallel": 2, "context_parallel_replicate": 2}, {"schedule": "1f1b", "num_stages_per_rank": 2, "zero_bubble": True}, False),
    ({"pipeline_parallel": 2, "context_parallel_replicate": 2, "expert_parallel": 2}, {"schedule": "zero_bubble_v"}, True),
    # split backward over FSDP-sharded stages: parameters stay unsharded between the input and the weight pass, one
    # reduce-scatter at the end of the pipeline step (D9D_PP_FSDP_SPLIT=require makes the fallback an error)
    ({"pipeline_parallel": 2, "context_parallel_shard": 2}, {"schedule": "1f1b", "num_stages_per_rank": 1, "zero_bubble": True}, False),
    ({"pipeline_parallel": 2, "context_parallel_shard": 2}, {"schedule": "zero_bubble_v"}, True),
]


@pytest.mark.dist
def test_pipeline_combinations_reproduce_the_single_process_run(tmp_path, monkeypatch):
    """Pipeline schedules combined with expert parallelism / FSDP / activation recomputation on meshes whose ranks all read the
    same samples (context parallel), so the loss trajectory can be compared with the single-process job step by step."""
    monkeypatch.setenv("D9D_PP_FSDP_SPLIT", "require")  # inherited by the spawned ranks
    layers = 4  # two stages per rank on two pipeline ranks need four layers
    jobs = []
    for moe in (False, True):
        _make_trainer(tmp_path / f"init{moe}", moe=moe, log=False, layers=layers).export(tmp_path / f"weights{moe}", load_checkpoint=False)
        _make_trainer(tmp_path / f"s{moe}", moe=moe, total_batch=8, micro=2, samples=16, source=tmp_path / f"weights{moe}", layers=layers).train()
    for i, (mesh_kwargs, schedule, moe) in enumerate(_PIPELINE_COMBINATIONS):
        jobs.append((str(tmp_path / f"job{i}"), mesh_kwargs, schedule, moe, str(tmp_path / f"weights{moe}"), True, layers, 16))
    run_distributed(_jobs_worker, 4, jobs)
    for i, (mesh_kwargs, schedule, moe) in enumerate(_PIPELINE_COMBINATIONS):
        _assert_same_trajectory(tmp_path / f"s{moe}", tmp_path / f"job{i}", (mesh_kwargs, schedule))


def test_resume_with_length_bucketed_data_lora_and_profiling(tmp_path):
    """Less common knobs in one job: a length-bucketing dataset whose shuffling state must be checkpointed, LoRA with
    adapter-only checkpoints, the profiler, no clipping, special step periods - interrupted and resumed exactly."""
    from tests.helpers_train import LMProvider, SFTTask, dense_params, trainer_config

    from d9d_b200.dataset import BufferSortedDataset, SyntheticTokenDataset, shard_dataset_data_parallel
    from d9d_b200.loop.auto import AutoLRSchedulerProvider, AutoOptimizerProvider
    from d9d_b200.loop.auto.auto_lr_scheduler import PiecewiseConfig
    from d9d_b200.loop.auto.auto_optimizer import AdamWOptimizerConfig
    from d9d_b200.loop.control import InitializeDatasetResult, InitializeModelStageResult
    from d9d_b200.loop.event.catalogue.train import EVENT_TRAIN_STEP_PRE
    from d9d_b200.loop.run import TrainingConfigurator
    from d9d_b200.model_state.mapper.compose import ModelStateMapperSequential
    from d9d_b200.peft import inject_peft_and_freeze
    from d9d_b200.peft.all import peft_method_from_config
    from d9d_b200.peft.lora.config import LoRAConfig

    class LoRAProvider(LMProvider):
        def initialize_model_stage(self, context):
            built = super().initialize_model_stage(context)
            method = peft_method_from_config(LoRAConfig.model_validate(
                {"kind": "lora", "module_name_pattern": r".*self_attn\.(q_proj|v_proj)", "params": {"r": 4, "alpha": 8, "dropout": 0.0}}))
            redirect = inject_peft_and_freeze(method, built.model)
            return InitializeModelStageResult(model=built.model, state_mapper=ModelStateMapperSequential([built.state_mapper, redirect]))

    def data_provider(context):
        base = SyntheticTokenDataset(64, 16, 128, seed=3, learnable=True)
        bucketed = BufferSortedDataset(base, buffer_size=16, pack_size=4, init_seed=5)
        return InitializeDatasetResult(dataset=shard_dataset_data_parallel(bucketed, context.dist_context), collator=SyntheticTokenDataset.collate)

    def make(tmp, profiling):
        cfg = trainer_config(tmp, total_batch=8, micro=4, ckpt_period=3, log_dir=tmp / "logs")
        cfg.model_stage_factory.checkpoint_only_trainable_parameters = True
        cfg.gradient_clipping.max_norm = None
        cfg.gc.period_steps = "disable"
        cfg.logging.period_steps = 1
        if profiling:
            from d9d_b200.loop.config import ProfilingConfig

            cfg.profiling = ProfilingConfig(enabled=True, traces_dir=tmp / "traces", period_steps=4, warmup_steps=1, active_steps=1)
        sched = PiecewiseConfig.model_validate({"name": "piecewise", "scheduler": {"initial_multiplier": 1.0, "phases": [
            {"mode": "rest", "target_multiplier": 0.5, "curve": {"type": "linear"}}]}})
        return TrainingConfigurator(
            mesh=__import__("d9d_b200.core.dist_context", fromlist=["DeviceMeshParameters"]).DeviceMeshParameters(), parameters=cfg,
            task_provider=lambda ctx: SFTTask(ctx.dist_context), model_provider=LoRAProvider(dense_params()), data_provider=data_provider,
            optimizer_provider=AutoOptimizerProvider(AdamWOptimizerConfig(lr=3e-3, weight_decay=0.0)),
            lr_scheduler_provider=AutoLRSchedulerProvider(sched)).configure()

    full = make(tmp_path / "full", profiling=True)
    full.train()
    traces = list((tmp_path / "full" / "traces").rglob("*trace.tar.gz"))
    assert traces and all(t.parent.name.startswith("step_") for t in traces)
    trainable = {n for n, p in full.state.tracked_modules.modules[0].named_parameters() if p.requires_grad}
    assert trainable and all("lora_" in n for n in trainable)

    class _Stop(Exception):
        pass

    def stop(ctx):
        if ctx.stepper.current_step == 5:
            raise _Stop

    part = make(tmp_path / "cut", profiling=False)
    part.state.event_bus.subscribe(EVENT_TRAIN_STEP_PRE, stop)
    with pytest.raises(_Stop):
        part.train()
    resumed = make(tmp_path / "cut", profiling=False)  # resumes from save-3: frozen base weights come from the deterministic init
    resumed.train()
    ref, _ = _read_losses(tmp_path / "full")
    got, _ = _read_losses(tmp_path / "cut")
    for step in range(3, 8):
        assert got[step] == ref[step], (step, got[step], ref[step])
    want = full.state.tracked_modules.modules[0].state_dict()
    for k, v in resumed.state.tracked_modules.modules[0].state_dict().items():
        torch.testing.assert_close(v, want[k], rtol=0, atol=0, msg=lambda m, k=k: f"{k}: {m}")


def test_fp8_linear_switch_of_the_model_stage_factory(tmp_path):
    """``model_stage_factory.fp8_linear`` re-classes the selected dense projections; the job trains (fp8 emulation on CPU)."""
    from d9d_b200.kernel.fp8 import Fp8Linear
    from d9d_b200.loop.config import Fp8LinearConfig

    trainer = _make_trainer(tmp_path)
    plain = trainer.state.tracked_modules.modules[0]
    assert not any(isinstance(m, Fp8Linear) for m in plain.modules())

    from d9d_b200.core.dist_context import DeviceMeshParameters
    from d9d_b200.loop.auto import AutoLRSchedulerProvider, AutoOptimizerProvider
    from d9d_b200.loop.auto.auto_lr_scheduler import PiecewiseConfig
    from d9d_b200.loop.auto.auto_optimizer import AdamWOptimizerConfig
    from d9d_b200.loop.run import TrainingConfigurator

    cfg = trainer_config(tmp_path / "fp8", log_dir=tmp_path / "fp8" / "logs")
    cfg.model_stage_factory.fp8_linear = Fp8LinearConfig(include=r"self_attn|mlp")
    sched = PiecewiseConfig.model_validate({"name": "piecewise", "scheduler": {"initial_multiplier": 1.0, "phases": [
        {"mode": "rest", "target_multiplier": 1.0, "curve": {"type": "linear"}}]}})
    # hidden 32 / intermediate 64 / q 32 / kv 16: every projection of the dense model is a multiple of 16
    fp8_trainer = TrainingConfigurator(
        mesh=DeviceMeshParameters(), parameters=cfg, task_provider=lambda ctx: SFTTask(ctx.dist_context),
        model_provider=LMProvider(dense_params(2)), data_provider=SyntheticDataProvider(num_samples=64),
        optimizer_provider=AutoOptimizerProvider(AdamWOptimizerConfig(lr=3e-3, weight_decay=0.0)),
        lr_scheduler_provider=AutoLRSchedulerProvider(sched)).configure()
    model = fp8_trainer.state.tracked_modules.modules[0]
    converted = [n for n, m in model.named_modules() if isinstance(m, Fp8Linear)]
    assert converted and all(("self_attn" in n or "mlp" in n) for n in converted) and not any("lm_head" in n for n in converted)
    assert set(model.state_dict()) == set(plain.state_dict())  # state-dict keys are unchanged
    fp8_trainer.train()
    losses, _ = _read_losses(tmp_path / "fp8")
    assert losses[max(losses)] < losses[min(losses)]
