"""Providers / task used by the Trainer tests (tiny causal-LM SFT on synthetic tokens)."""

from __future__ import annotations

import torch

from d9d_b200.dataset import SyntheticTokenDataset, shard_dataset_data_parallel
from d9d_b200.loop.control import (
    BuildForwardInputsContext,
    BuildForwardInputsResult,
    ComputeLossContext,
    ComputeLossResult,
    CreateMetricsContext,
    CreateMetricsResult,
    DatasetProvider,
    InitializeDatasetContext,
    InitializeDatasetResult,
    InitializeModelStageContext,
    InitializeModelStageResult,
    ModelProvider,
    ParallelizeModelStageContext,
    PrepareExportModelStageContext,
    PrepareExportModelStageResult,
    TrainTask,
    UpdateMetricsContext,
)
from d9d_b200.metric.impl.aggregation import SumMetric
from d9d_b200.model_state.mapper.adapters import identity_mapper_from_module
from d9d_b200.module.block.head import LM_IGNORE_INDEX
from d9d_b200.module.block.hidden_states_aggregator import HiddenStatesAggregationMode

VOCAB = {"regular": 100, "special": 28}


def dense_params(layers: int = 2):
    from d9d_b200.module.model.qwen3_dense import Qwen3DenseForCausalLMParameters, Qwen3DenseLayerParameters, Qwen3DenseParameters

    return Qwen3DenseForCausalLMParameters(model=Qwen3DenseParameters(
        layer=Qwen3DenseLayerParameters(hidden_size=32, intermediate_size=64, num_attention_heads=4, num_key_value_heads=2,
                                        rms_norm_eps=1e-6, head_dim=8),
        num_hidden_layers=layers, rope_base=10000, max_position_ids=64, split_vocab_size=VOCAB,
        split_vocab_order=["regular", "special"], pipeline_num_virtual_layers_post=0))


def moe_params(layers: int = 2):
    from d9d_b200.module.model.qwen3_moe import Qwen3MoEForCausalLMParameters, Qwen3MoELayerParameters, Qwen3MoEParameters

    return Qwen3MoEForCausalLMParameters(model=Qwen3MoEParameters(
        layer=Qwen3MoELayerParameters(hidden_size=32, intermediate_size=16, num_experts=4, experts_top_k=2, num_attention_heads=4,
                                      num_key_value_heads=2, rms_norm_eps=1e-6, head_dim=8),
        num_hidden_layers=layers, rope_base=10000, max_position_ids=64, split_vocab_size=VOCAB,
        split_vocab_order=["regular", "special"]))


class SyntheticDataProvider(DatasetProvider):
    def __init__(self, num_samples: int = 64, seq_len: int = 16, seed: int = 3):
        self._n, self._s, self._seed = num_samples, seq_len, seed

    def __call__(self, context: InitializeDatasetContext) -> InitializeDatasetResult:
        data = SyntheticTokenDataset(self._n, self._s, sum(VOCAB.values()), seed=self._seed, learnable=True)
        return InitializeDatasetResult(dataset=shard_dataset_data_parallel(data, context.dist_context), collator=SyntheticTokenDataset.collate)


class LMProvider(ModelProvider):
    def __init__(self, params, moe: bool = False, dtype: torch.dtype = torch.float32, activation_checkpointing: bool = False):
        self._params, self._moe, self._dtype, self._recompute = params, moe, dtype, activation_checkpointing

    def initialize_model_stage(self, context: InitializeModelStageContext) -> InitializeModelStageResult:
        if self._moe:
            from d9d_b200.module.model.qwen3_moe import Qwen3MoEForCausalLM as Cls
        else:
            from d9d_b200.module.model.qwen3_dense import Qwen3DenseForCausalLM as Cls
        model = Cls(self._params, context.stage, HiddenStatesAggregationMode.no, self._recompute).to(self._dtype)
        return InitializeModelStageResult(model=model, state_mapper=identity_mapper_from_module(model))

    def parallelize_model_stage(self, context: ParallelizeModelStageContext) -> None:
        if self._moe:
            from d9d_b200.module.parallelism.model.qwen3_moe import parallelize_qwen3_moe_for_causal_lm as fn
        else:
            from d9d_b200.module.parallelism.model.qwen3_dense import parallelize_qwen3_dense_for_causal_lm as fn
        fn(context.dist_context, context.model, context.stage)

    def prepare_export_model_stage(self, context: PrepareExportModelStageContext) -> PrepareExportModelStageResult:
        return PrepareExportModelStageResult(state_mapper=identity_mapper_from_module(context.model))

    def dump_hparams(self):
        return self._params.model_dump(mode="json")


class SFTTask(TrainTask):
    def __init__(self, dist_context=None):
        self._ctx = dist_context

    def build_forward_inputs(self, ctx: BuildForwardInputsContext) -> BuildForwardInputsResult:
        from d9d_b200.dataset import shard_batch_along_sequence

        batch = ctx.batch if self._ctx is None else shard_batch_along_sequence(ctx.batch, self._ctx)
        ctx.state["labels"] = batch["labels"]
        return BuildForwardInputsResult(inputs={"input_ids": batch["input_ids"]},
                                        kwargs={"labels": batch["labels"], "position_ids": batch["position_ids"]})

    def create_metrics(self, ctx: CreateMetricsContext) -> CreateMetricsResult:
        return CreateMetricsResult(metrics={"num_tokens": SumMetric()})

    def update_metrics(self, ctx: UpdateMetricsContext) -> None:
        ctx.metrics["num_tokens"].update(ctx.state["num_tokens"])

    def compute_loss(self, ctx: ComputeLossContext) -> ComputeLossResult:
        n = (ctx.state["labels"] != LM_IGNORE_INDEX).sum()
        ctx.state["num_tokens"] = n
        return ComputeLossResult(loss=ctx.pipeline_results["logps"].sum() / n.clamp_min(1), loss_weight=n / 1000)


def trainer_config(tmp, total_batch=8, micro=4, schedule=None, ckpt_period="disable", log_dir=None, source=None,
                   async_save=False):
    from d9d_b200.loop.config import TrainerConfig

    return TrainerConfig.model_validate({
        "run": {"name": "t", "description": None, "hparams": {}},
        "batching": {"global_batch_size": total_batch, "microbatch_size": micro},
        "data_loading": {"num_workers": 0, "pin_memory": False, "persistent_workers": False},
        "logging": {"period_steps": 2, "tracker": {"provider": "jsonl", "directory": str(log_dir)} if log_dir else {"provider": "null"}},
        "pipelining": {"schedule": schedule or {"schedule": "gpipe"}},
        "model_stage_factory": {"source_checkpoint": str(source) if source else None, "checkpoint_only_trainable_parameters": False},
        "determinism": {"base_seed": 11},
        "gc": {"period_steps": 4},
        "checkpointing": {"save_dir": str(tmp / "ckpt"), "period_steps": ckpt_period, "num_to_keep": 2, "async_save": async_save},
        "gradient_clipping": {"max_norm": 1.0, "log_total_steps": 2},
        "profiling": None,
        "gradient_manager": {"grad_dtype": "float32", "bucket_size_mb": 1},
        "timeout": {"init_timeout": 600, "step_timeout": 300},
    })
