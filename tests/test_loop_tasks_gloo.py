"""Loops beyond causal-LM training on multi-rank meshes: the inference loop over a pipeline (results collected on the last
stage of every data-parallel replica) and a classification job (pooled head, label metrics) over pipeline x data parallel."""

import json

import pytest
import torch

from tests.dist_utils import run_distributed

pytestmark = pytest.mark.dist


# ------------------------------------------------------------------------------------------------ inference
def _inference(tmp, mesh_kwargs, source):
    from pathlib import Path

    from tests.helpers_train import LMProvider, SyntheticDataProvider, dense_params

    from d9d_b200.core.dist_context import DeviceMeshParameters
    from d9d_b200.loop.config import InferenceConfig
    from d9d_b200.loop.control import BuildForwardInputsResult, InferenceTask
    from d9d_b200.loop.run import InferenceConfigurator

    class Perplexity(InferenceTask):
        def __init__(self):
            self.nll, self.tokens, self.calls = 0.0, 0, 0

        def build_forward_inputs(self, ctx):
            ctx.state["labels"] = ctx.batch["labels"]
            return BuildForwardInputsResult(inputs={"input_ids": ctx.batch["input_ids"]},
                                            kwargs={"labels": ctx.batch["labels"], "position_ids": ctx.batch["position_ids"]})

        def process_outputs(self, ctx):
            self.calls += 1
            self.nll += float(ctx.pipeline_results["logps"].sum())
            self.tokens += int((ctx.state["labels"] != -100).sum())

    task = Perplexity()
    config = InferenceConfig.model_validate({
        "batching": {"global_batch_size": 8, "microbatch_size": 2},
        "data_loading": {"num_workers": 0, "pin_memory": False, "persistent_workers": False},
        "model_stage_factory": {"source_checkpoint": str(source), "checkpoint_only_trainable_parameters": False},
        "determinism": {"base_seed": 0}, "gc": {"period_steps": "disable"},
        "checkpointing": {"save_dir": str(Path(tmp) / "progress"), "period_steps": "disable", "num_to_keep": None}, "profiling": None})
    job = InferenceConfigurator(mesh=DeviceMeshParameters(**mesh_kwargs), parameters=config, task_provider=lambda ctx: task,
                                model_provider=LMProvider(dense_params()), data_provider=SyntheticDataProvider(num_samples=32)).configure()
    job.infer()
    assert not any(p.requires_grad and p.grad is not None for m in job.state.tracked_modules.modules for p in m.parameters())
    return task, job


def _inference_worker(rank, world, tmp, mesh_kwargs, source):
    from pathlib import Path

    task, job = _inference(tmp, mesh_kwargs, source)
    (Path(tmp) / f"rank{rank}.json").write_text(json.dumps({"nll": task.nll, "tokens": task.tokens, "calls": task.calls}))


def test_inference_loop_over_pipeline_and_data_parallel(tmp_path):
    from tests.helpers_train import LMProvider, SFTTask, SyntheticDataProvider, dense_params, trainer_config
    from tests.test_trainer import _make_trainer

    del LMProvider, SFTTask, SyntheticDataProvider, dense_params, trainer_config
    _make_trainer(tmp_path / "init", log=False).export(tmp_path / "weights", load_checkpoint=False)
    single, _ = _inference(tmp_path / "single", {}, tmp_path / "weights")
    assert single.tokens == 32 * 16 and single.calls == 16  # 32 samples, one call per microbatch of 2

    (tmp_path / "dist").mkdir()
    run_distributed(_inference_worker, 4, str(tmp_path / "dist"), {"pipeline_parallel": 2, "data_parallel_replicate": 2},
                    str(tmp_path / "weights"))
    per_rank = [json.loads((tmp_path / "dist" / f"rank{r}.json").read_text()) for r in range(4)]
    last_stage = [r for r in per_rank if r["calls"] > 0]
    assert len(last_stage) == 2  # only the ranks holding the last pipeline stage see outputs, one per data-parallel replica
    assert sum(r["tokens"] for r in last_stage) == single.tokens
    assert abs(sum(r["nll"] for r in last_stage) - single.nll) < 1e-3 * abs(single.nll)


# ------------------------------------------------------------------------------------------- classification
def _classification_worker(rank, world, tmp, mesh_kwargs, head_only=False):
    from pathlib import Path

    from tests.helpers_train import trainer_config

    from d9d_b200.core.dist_context import DeviceMeshParameters
    from d9d_b200.dataset import TokenPoolingType, shard_dataset_data_parallel, token_pooling_mask_from_attention_mask
    from d9d_b200.loop.auto import AutoLRSchedulerProvider, AutoOptimizerProvider
    from d9d_b200.loop.auto.auto_lr_scheduler import PiecewiseConfig
    from d9d_b200.loop.auto.auto_optimizer import AdamWOptimizerConfig
    from d9d_b200.loop.control import (BuildForwardInputsResult, ComputeLossResult, CreateMetricsResult, InitializeDatasetResult,
                                       InitializeModelStageResult, ModelProvider, PrepareExportModelStageResult, TrainTask)
    from d9d_b200.loop.run import TrainingConfigurator
    from d9d_b200.metric.impl.classification.confusion_matrix import confusion_matrix_metric
    from d9d_b200.model_state.mapper.adapters import identity_mapper_from_module
    from d9d_b200.module.block.hidden_states_aggregator import HiddenStatesAggregationMode
    from d9d_b200.module.model.qwen3_dense import (Qwen3DenseForClassification, Qwen3DenseForClassificationParameters,
                                                   Qwen3DenseLayerParameters, Qwen3DenseParameters)
    from d9d_b200.module.parallelism.model.qwen3_dense import parallelize_qwen3_dense_for_classification

    params = Qwen3DenseForClassificationParameters(model=Qwen3DenseParameters(
        layer=Qwen3DenseLayerParameters(hidden_size=32, intermediate_size=64, num_attention_heads=4, num_key_value_heads=2,
                                        rms_norm_eps=1e-6, head_dim=8),
        num_hidden_layers=2, rope_base=10000, max_position_ids=64, split_vocab_size={"text": 16}, split_vocab_order=["text"]),
        num_labels=2, classifier_dropout=0.0)

    class Provider(ModelProvider):
        def initialize_model_stage(self, context):
            model = Qwen3DenseForClassification(params, context.stage, HiddenStatesAggregationMode.no, False)
            if head_only:  # tune the classification head only: the first pipeline stage ends up without trainable parameters
                from d9d_b200.peft import inject_peft_and_freeze
                from d9d_b200.peft.all import peft_method_from_config
                from d9d_b200.peft.full_tune.config import FullTuneConfig

                inject_peft_and_freeze(peft_method_from_config(FullTuneConfig.model_validate(
                    {"kind": "full_tune", "module_name_pattern": "cls_head.*"})), model)
            return InitializeModelStageResult(model=model, state_mapper=identity_mapper_from_module(model))

        def parallelize_model_stage(self, context):
            parallelize_qwen3_dense_for_classification(context.dist_context, context.model, context.stage)

        def prepare_export_model_stage(self, context):
            return PrepareExportModelStageResult(state_mapper=identity_mapper_from_module(context.model))

    class Majority(torch.utils.data.Dataset):
        def __init__(self):
            g = torch.Generator().manual_seed(0)
            self.rows = torch.randint(1, 4, (128, 12), generator=g)

        def __len__(self):
            return len(self.rows)

        def __getitem__(self, i):
            row = self.rows[i]
            return {"input_ids": row, "label": ((row == 1).sum() > (row == 2).sum()).long()}

    def collate(batch):
        ids = torch.stack([b["input_ids"] for b in batch])
        mask = token_pooling_mask_from_attention_mask(torch.ones_like(ids), TokenPoolingType.last)
        return {"input_ids": ids, "position_ids": torch.arange(ids.shape[1]).expand_as(ids), "pooling_mask": mask,
                "labels": torch.stack([b["label"] for b in batch])}

    class Task(TrainTask):
        def build_forward_inputs(self, ctx):
            ctx.state["labels"] = ctx.batch["labels"]
            return BuildForwardInputsResult(inputs={"input_ids": ctx.batch["input_ids"]},
                                            kwargs={"position_ids": ctx.batch["position_ids"], "pooling_mask": ctx.batch["pooling_mask"]})

        def compute_loss(self, ctx):
            scores, labels = ctx.pipeline_results["scores"], ctx.state["labels"]
            assert scores.shape == (labels.shape[0], 2)
            ctx.state["scores"] = scores.detach()
            return ComputeLossResult(loss=torch.nn.functional.cross_entropy(scores, labels), loss_weight=torch.tensor(float(len(labels))))

        def create_metrics(self, ctx):
            return CreateMetricsResult(metrics={"accuracy": confusion_matrix_metric().multiclass(2, top_k=1).with_accuracy().build()})

        def update_metrics(self, ctx):
            ctx.metrics["accuracy"].update(ctx.state["scores"], ctx.state["labels"])

    tmp = Path(tmp)
    schedule = PiecewiseConfig.model_validate({"name": "piecewise", "scheduler": {"initial_multiplier": 1.0, "phases": [
        {"mode": "rest", "target_multiplier": 0.2, "curve": {"type": "cosine"}}]}})
    trainer = TrainingConfigurator(
        mesh=DeviceMeshParameters(**mesh_kwargs),
        parameters=trainer_config(tmp, total_batch=16, micro=4, schedule={"schedule": "1f1b", "num_stages_per_rank": 1, "zero_bubble": True},
                                  log_dir=tmp / "logs"),
        task_provider=lambda ctx: Task(), model_provider=Provider(),
        data_provider=lambda ctx: InitializeDatasetResult(dataset=shard_dataset_data_parallel(Majority(), ctx.dist_context), collator=collate),
        optimizer_provider=AutoOptimizerProvider(AdamWOptimizerConfig(lr=5e-3, weight_decay=0.0)),
        lr_scheduler_provider=AutoLRSchedulerProvider(schedule)).configure()
    trainer.train()
    if head_only:
        trainable = [n for m in trainer.state.tracked_modules.modules for n, p in m.named_parameters() if p.requires_grad]
        assert all(n.startswith("cls_head") for n in trainable)
        assert bool(trainable) == trainer.state.tracked_modules.modules[0]._stage.is_current_stage_last  # noqa: SLF001


def test_head_only_tuning_with_a_fully_frozen_pipeline_stage(tmp_path):
    """Only the classification head (last stage) is trainable: the other stage has no trainable parameter at all, which the
    optimizer factory, gradient synchronisation and clipping have to cope with."""
    run_distributed(_classification_worker, 2, str(tmp_path), {"pipeline_parallel": 2}, True)
    records = [json.loads(line) for line in next((tmp_path / "logs").glob("*.jsonl")).read_text().splitlines()]
    losses = [r["value"] for r in records if r.get("name") == "loss"]
    assert len(losses) == 8 and all(v == v for v in losses)


def test_classification_job_over_pipeline_and_data_parallel(tmp_path):
    run_distributed(_classification_worker, 4, str(tmp_path), {"pipeline_parallel": 2, "data_parallel_replicate": 2})
    records = [json.loads(line) for line in next((tmp_path / "logs").glob("*.jsonl")).read_text().splitlines()]
    losses = [r["value"] for r in records if r.get("name") == "loss"]
    accuracy = [r["value"] for r in records if r.get("name") == "accuracy"]
    assert len(losses) == 8 and all(v == v for v in losses) and losses[-1] < losses[0]
    assert accuracy and 0.0 <= accuracy[-1] <= 1.0


# ------------------------------------------------------------------------------------- resumable evaluation

def _perplexity_job(tmp, task, ckpt_period):
    from d9d_b200.core.dist_context import DeviceMeshParameters
    from d9d_b200.loop.config import InferenceConfig
    from d9d_b200.loop.run import InferenceConfigurator
    from tests.helpers_train import LMProvider, SyntheticDataProvider, dense_params
    config = InferenceConfig.model_validate({
        "batching": {"global_batch_size": 8, "microbatch_size": 4},
        "data_loading": {"num_workers": 0, "pin_memory": False, "persistent_workers": False},
        "model_stage_factory": {"source_checkpoint": None, "checkpoint_only_trainable_parameters": False},
        "determinism": {"base_seed": 0}, "gc": {"period_steps": "disable"},
        "checkpointing": {"save_dir": str(tmp / "progress"), "period_steps": ckpt_period, "num_to_keep": None}, "profiling": None})
    return InferenceConfigurator(mesh=DeviceMeshParameters(), parameters=config, task_provider=lambda ctx: task,
                                 model_provider=LMProvider(dense_params()), data_provider=SyntheticDataProvider(num_samples=32)).configure()

def test_interrupted_inference_resumes_with_its_running_sums(tmp_path):
    """The perplexity task checkpoints its sums (per rank): an evaluation killed after two batches and restarted reports the
    same totals as an uninterrupted one."""
    from d9d_b200.loop.event.catalogue.inference import EVENT_INFERENCE_STEP_PRE
    from d9d_b200.recipes import CausalLMPerplexityTask
    full = CausalLMPerplexityTask()
    _perplexity_job(tmp_path / "full", full, "disable").infer()
    class Stop(Exception): pass
    part = CausalLMPerplexityTask()
    job = _perplexity_job(tmp_path / "cut", part, 1)
    def stop(ctx):
        if ctx.stepper.current_step == 2: raise Stop
    job.state.event_bus.subscribe(EVENT_INFERENCE_STEP_PRE, stop)
    with pytest.raises(Stop): job.infer()
    resumed = CausalLMPerplexityTask()
    _perplexity_job(tmp_path / "cut", resumed, 1).infer()
    assert resumed.num_tokens == full.num_tokens and abs(resumed.nll_sum - full.nll_sum) < 1e-6 * abs(full.nll_sum)
