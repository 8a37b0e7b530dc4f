"""Perplexity of a Qwen3-MoE checkpoint over a dataset (forward-only pipeline schedule).

    torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node 8 calculate_perplexity.py calculate_perplexity.json
    python calculate_perplexity.py calculate_perplexity.json --single
"""

from __future__ import annotations

import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))  # run in place without installing the package

import argparse
import math

import torch
import torch.distributed as dist
from pydantic import BaseModel

from d9d_b200.core.dist_context import DeviceMeshParameters
from d9d_b200.loop.config import InferenceConfig
from d9d_b200.loop.run import InferenceConfigurator
from d9d_b200.recipes import CausalLMPerplexityTask, Qwen3MoEModelProvider, Qwen3MoEModelProviderConfig, SyntheticDataProvider
from pretrain import SyntheticData, TextDataConfig, TextDataProvider


class ProjectConfig(BaseModel):
    mesh: DeviceMeshParameters
    data: TextDataConfig | SyntheticData
    model_provider: Qwen3MoEModelProviderConfig
    inference: InferenceConfig


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("config", nargs="?", default=str(Path(__file__).with_name("calculate_perplexity.json")))
    ap.add_argument("--single", action="store_true")
    args = ap.parse_args()
    config = ProjectConfig.model_validate_json(Path(args.config).read_text(encoding="utf-8"))
    tasks: list[CausalLMPerplexityTask] = []

    def task_provider(ctx):  # noqa: ANN001, ANN202  rank-aware: the running sums are checkpointed per rank
        tasks.append(CausalLMPerplexityTask(ctx.dist_context))
        return tasks[-1]

    data_provider = SyntheticDataProvider(config.data) if config.data.kind == "synthetic" else TextDataProvider(config.data)
    job = InferenceConfigurator(
        mesh=DeviceMeshParameters() if args.single else config.mesh,
        parameters=config.inference,
        task_provider=task_provider,
        model_provider=Qwen3MoEModelProvider(config.model_provider),
        data_provider=data_provider,
    ).configure()
    job.infer()
    task = tasks[0]

    # every data-parallel replica saw a different shard: sum the statistics before reporting
    stats = torch.tensor([task.nll_sum, float(task.num_tokens)], dtype=torch.float64)
    if dist.is_initialized():
        stats = stats.to(job.state.dist_context.current_device)
        dist.all_reduce(stats)
    if job.state.dist_context.is_main_process and stats[1] > 0:
        print(f"tokens={int(stats[1])} nll/token={float(stats[0] / stats[1]):.4f} perplexity={math.exp(float(stats[0] / stats[1])):.3f}")


if __name__ == "__main__":
    main()
