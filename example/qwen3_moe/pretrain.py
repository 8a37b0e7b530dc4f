"""Qwen3-MoE pre-training example.

    # 8 GPUs, PP=4 x DP=2, EP=2 (the mesh in pretrain.json)
    torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node 8 pretrain.py pretrain.json
    # single GPU
    python pretrain.py pretrain.json --single

Data: ``data.kind = "synthetic"`` trains on random tokens (no network needed); ``"text"`` tokenises a HuggingFace
dataset column with a ``tokenizers`` tokenizer file and batches by length through ``BufferSortedDataset``.
"""

from __future__ import annotations

import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))  # run in place without installing the package

import argparse
from collections.abc import Sequence
from typing import Any, Literal

import torch
from pydantic import BaseModel
from torch.utils.data import Dataset

from d9d_b200.core.dist_context import DeviceMeshParameters
from d9d_b200.dataset import BufferSortedDataset, pad_stack_1d, shard_dataset_data_parallel
from d9d_b200.loop.auto import AutoLRSchedulerConfig, AutoLRSchedulerProvider, AutoOptimizerConfig, AutoOptimizerProvider
from d9d_b200.loop.config import TrainerConfig
from d9d_b200.loop.control import DatasetProvider, InitializeDatasetContext, InitializeDatasetResult
from d9d_b200.loop.run import TrainingConfigurator
from d9d_b200.module.block.head import LM_IGNORE_INDEX
from d9d_b200.recipes import (
    CausalLMTask,
    Qwen3MoEModelProvider,
    Qwen3MoEModelProviderConfig,
    SyntheticDataConfig,
    SyntheticDataProvider,
)


class TextDataConfig(BaseModel):
    kind: Literal["text"] = "text"
    dataset: str
    split: str
    text_column: str
    use_samples: int
    shuffle_seed: int
    tokenizer: str
    presort_buffer_size: int
    num_proc: int


class SyntheticData(SyntheticDataConfig):
    kind: Literal["synthetic"] = "synthetic"


class ProjectConfig(BaseModel):
    mesh: DeviceMeshParameters
    data: TextDataConfig | SyntheticData
    model_provider: Qwen3MoEModelProviderConfig
    trainer: TrainerConfig
    optimizer: AutoOptimizerConfig
    lr_scheduler: AutoLRSchedulerConfig
    export_to: Path


class TokenisedTextDataset(Dataset):
    """Tokenises on access; ``sort_key`` (token count) lets ``BufferSortedDataset`` build low-padding batches."""

    def __init__(self, rows: Any, tokenizer: Any, text_column: str):
        self._rows, self._tokenizer, self._column = rows, tokenizer, text_column

    def __len__(self) -> int:
        return len(self._rows)

    def sort_key(self, index: int) -> int:
        return self._rows[index]["token_counts"]

    def __getitem__(self, index: int) -> dict[str, torch.Tensor]:
        tokens = torch.tensor(self._tokenizer.encode(self._rows[index][self._column]).ids, dtype=torch.long)
        # the models do not shift labels: inputs are tokens[:-1], targets tokens[1:]
        return {"input_ids": tokens[:-1], "labels": tokens[1:], "position_ids": torch.arange(tokens.numel() - 1)}

    @staticmethod
    def collate(batch: Sequence[dict[str, torch.Tensor]]) -> dict[str, torch.Tensor]:
        return {
            "input_ids": pad_stack_1d([b["input_ids"] for b in batch], pad_value=0),
            "labels": pad_stack_1d([b["labels"] for b in batch], pad_value=LM_IGNORE_INDEX),
            "position_ids": pad_stack_1d([b["position_ids"] for b in batch], pad_value=0),
        }


class TextDataProvider(DatasetProvider):
    def __init__(self, config: TextDataConfig):
        self._config = config

    def __call__(self, context: InitializeDatasetContext) -> InitializeDatasetResult:
        import datasets
        from tokenizers import Tokenizer

        c = self._config
        tokenizer = Tokenizer.from_file(c.tokenizer)
        with context.dist_context.main_process_first():  # rank 0 fills the HF cache, the others read it
            rows = (datasets.load_dataset(c.dataset, split=c.split).take(c.use_samples).shuffle(c.shuffle_seed)
                    .map(lambda item: {"token_counts": len(tokenizer.encode(item[c.text_column]).ids)}, num_proc=c.num_proc))
        data = BufferSortedDataset(TokenisedTextDataset(rows, tokenizer, c.text_column), buffer_size=c.presort_buffer_size,
                                   pack_size=context.batch_maths.global_batch_size, init_seed=c.shuffle_seed)
        return InitializeDatasetResult(dataset=shard_dataset_data_parallel(data, context.dist_context),
                                       collator=TokenisedTextDataset.collate)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("config", nargs="?", default=str(Path(__file__).with_name("pretrain.json")))
    ap.add_argument("--single", action="store_true", help="ignore the mesh in the config and run on one device")
    args = ap.parse_args()

    config = ProjectConfig.model_validate_json(Path(args.config).read_text(encoding="utf-8"))
    mesh = DeviceMeshParameters() if args.single else config.mesh
    data_provider = SyntheticDataProvider(config.data) if config.data.kind == "synthetic" else TextDataProvider(config.data)

    trainer = TrainingConfigurator(
        mesh=mesh,
        parameters=config.trainer,
        task_provider=lambda ctx: CausalLMTask(ctx.dist_context),  # context-parallel meshes shard the sequence in the task
        model_provider=Qwen3MoEModelProvider(config.model_provider),
        data_provider=data_provider,
        optimizer_provider=AutoOptimizerProvider(config.optimizer),
        lr_scheduler_provider=AutoLRSchedulerProvider(config.lr_scheduler),
    ).configure()
    trainer.train()
    trainer.export(config.export_to, load_checkpoint=False)


if __name__ == "__main__":
    main()
