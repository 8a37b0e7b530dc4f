"""Sequence classification with LoRA on a Qwen3-dense backbone.

    python finetune.py finetune.json                 # one device
    torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node 8 finetune.py finetune.json

Shows the pieces the pre-training example does not touch: a classification head over pooled tokens, parameter-efficient
fine-tuning (LoRA on the attention projections, full tuning of the head) with the checkpoint-redirecting state mapper,
classification metrics, adapter-only job checkpoints and an export with the adapters merged back into plain weights.
The task is synthetic (no network needed): does token 1 occur more often than token 2 in the sequence?
"""

from __future__ import annotations

import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))  # run in place without installing the package

import argparse

import torch
from pydantic import BaseModel
from torch.utils.data import Dataset

from d9d_b200.core.dist_context import DeviceMeshParameters
from d9d_b200.dataset import TokenPoolingType, pad_stack_1d, shard_dataset_data_parallel, token_pooling_mask_from_attention_mask
from d9d_b200.loop.auto import AutoLRSchedulerConfig, AutoLRSchedulerProvider, AutoOptimizerConfig, AutoOptimizerProvider
from d9d_b200.loop.config import TrainerConfig
from d9d_b200.loop.control import (
    BuildForwardInputsContext,
    BuildForwardInputsResult,
    ComputeLossContext,
    ComputeLossResult,
    CreateMetricsContext,
    CreateMetricsResult,
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
from d9d_b200.loop.run import TrainingConfigurator
from d9d_b200.metric.impl.classification.confusion_matrix import confusion_matrix_metric
from d9d_b200.model_state.mapper.compose import ModelStateMapperSequential
from d9d_b200.module.block.hidden_states_aggregator import HiddenStatesAggregationMode
from d9d_b200.module.model.qwen3_dense import (
    Qwen3DenseForClassification,
    Qwen3DenseForClassificationParameters,
    mapper_from_huggingface_qwen3_dense,
    mapper_to_huggingface_qwen3_dense_for_classification,
)
from d9d_b200.module.parallelism.model.qwen3_dense import parallelize_qwen3_dense_for_classification
from d9d_b200.peft import inject_peft_and_freeze, merge_peft
from d9d_b200.peft.all import AnyPeftConfig, peft_method_from_config


class MajorityDataset(Dataset):
    """Random variable-length token sequences labelled by whether token ``1`` outnumbers token ``2``."""

    def __init__(self, num_samples: int, max_len: int, vocab_size: int, seed: int):
        generator = torch.Generator().manual_seed(seed)
        lengths = torch.randint(max_len // 2, max_len + 1, (num_samples,), generator=generator)
        self._rows = [torch.randint(1, min(vocab_size, 5), (int(n),), generator=generator) for n in lengths]

    def __len__(self) -> int:
        return len(self._rows)

    def __getitem__(self, index: int) -> dict[str, torch.Tensor]:
        tokens = self._rows[index]
        return {"input_ids": tokens, "label": ((tokens == 1).sum() > (tokens == 2).sum()).long()}

    @staticmethod
    def collate(batch: list[dict[str, torch.Tensor]]) -> dict[str, torch.Tensor]:
        ids = pad_stack_1d([b["input_ids"] for b in batch], pad_value=0)
        attention = (ids != 0).long()
        return {"input_ids": ids, "position_ids": (attention.cumsum(1) - 1).clamp_min(0),
                "pooling_mask": token_pooling_mask_from_attention_mask(attention, TokenPoolingType.last),
                "labels": torch.stack([b["label"] for b in batch])}


class DataConfig(BaseModel):
    num_samples: int
    max_len: int
    seed: int


class ProjectConfig(BaseModel):
    mesh: DeviceMeshParameters
    data: DataConfig
    model: Qwen3DenseForClassificationParameters
    pretrained_backbone: Path | None  # HuggingFace-format Qwen3 backbone (sharded safetensors) or null for random weights
    peft: AnyPeftConfig
    trainer: TrainerConfig
    optimizer: AutoOptimizerConfig
    lr_scheduler: AutoLRSchedulerConfig
    export_to: Path


class ClassifierProvider(ModelProvider):
    def __init__(self, config: ProjectConfig):
        self._config = config
        self._peft = peft_method_from_config(config.peft)

    def initialize_model_stage(self, context: InitializeModelStageContext) -> InitializeModelStageResult:
        model = Qwen3DenseForClassification(self._config.model, context.stage, HiddenStatesAggregationMode.no, False)
        # adapters change the module tree; the returned mapper redirects stock checkpoint keys (x.weight -> x.base.weight)
        redirect = inject_peft_and_freeze(self._peft, model)
        from_hf = mapper_from_huggingface_qwen3_dense(self._config.model.model)  # bare backbone: the head is new
        return InitializeModelStageResult(model=model, state_mapper=ModelStateMapperSequential([_under_model_prefix(from_hf), redirect]))

    def parallelize_model_stage(self, context: ParallelizeModelStageContext) -> None:
        parallelize_qwen3_dense_for_classification(context.dist_context, context.model, context.stage)

    def prepare_export_model_stage(self, context: PrepareExportModelStageContext) -> PrepareExportModelStageResult:
        merge_peft(self._peft, context.model)  # fold the adapters in: the export is a plain HuggingFace classifier
        return PrepareExportModelStageResult(state_mapper=mapper_to_huggingface_qwen3_dense_for_classification(self._config.model))

    def dump_hparams(self) -> dict:
        return {"model": self._config.model.model_dump(mode="json"), "peft": self._config.peft.model_dump(mode="json")}


def _under_model_prefix(mapper):  # noqa: ANN001, ANN202
    from d9d_b200.model_state.mapper.compose import ModelStateMapperPrefixScope

    return ModelStateMapperPrefixScope(mapper, source_prefix="", target_prefix="model.")


class ClassificationTask(TrainTask):
    def build_forward_inputs(self, ctx: BuildForwardInputsContext) -> BuildForwardInputsResult:
        ctx.state["labels"] = ctx.batch["labels"]
        return BuildForwardInputsResult(inputs={"input_ids": ctx.batch["input_ids"]},
                                        kwargs={"position_ids": ctx.batch["position_ids"], "pooling_mask": ctx.batch["pooling_mask"]})

    def compute_loss(self, ctx: ComputeLossContext) -> ComputeLossResult:
        scores, labels = ctx.pipeline_results["scores"], ctx.state["labels"]
        ctx.state["scores"] = scores.detach()
        count = torch.tensor(float(labels.numel()), device=scores.device)
        return ComputeLossResult(loss=torch.nn.functional.cross_entropy(scores, labels), loss_weight=count)

    def create_metrics(self, ctx: CreateMetricsContext) -> CreateMetricsResult:
        return CreateMetricsResult(metrics={
            "accuracy": confusion_matrix_metric().multiclass(num_classes=2, top_k=1).with_accuracy().build(),
            "f1_macro": confusion_matrix_metric().multiclass(num_classes=2).with_f1().macro().build()})

    def update_metrics(self, ctx: UpdateMetricsContext) -> None:
        for metric in ctx.metrics.values():
            metric.update(ctx.state["scores"], ctx.state["labels"])


def main(argv: list[str] | None = None) -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("config", nargs="?", default=str(Path(__file__).with_name("finetune.json")))
    args = ap.parse_args(argv)
    config = ProjectConfig.model_validate_json(Path(args.config).read_text(encoding="utf-8"))
    if config.pretrained_backbone is not None:
        config.trainer.model_stage_factory.source_checkpoint = config.pretrained_backbone

    def data_provider(context: InitializeDatasetContext) -> InitializeDatasetResult:
        data = MajorityDataset(config.data.num_samples, config.data.max_len, config.model.model.vocab_size, config.data.seed)
        return InitializeDatasetResult(dataset=shard_dataset_data_parallel(data, context.dist_context), collator=MajorityDataset.collate)

    trainer = TrainingConfigurator(
        mesh=config.mesh, parameters=config.trainer, task_provider=lambda ctx: ClassificationTask(),
        model_provider=ClassifierProvider(config), data_provider=data_provider,
        optimizer_provider=AutoOptimizerProvider(config.optimizer), lr_scheduler_provider=AutoLRSchedulerProvider(config.lr_scheduler),
    ).configure()
    trainer.train()
    trainer.export(config.export_to, load_checkpoint=False)


if __name__ == "__main__":
    main()
