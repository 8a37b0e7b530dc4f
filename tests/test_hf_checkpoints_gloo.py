"""HuggingFace-format checkpoints through the training loop on a pipeline x FSDP x expert-parallel mesh: every pipeline stage
streams its part of a whole-model ``Qwen3MoeForCausalLM`` checkpoint into sharded parameters (fused 3-D expert tensors ->
expert-sharded ``GroupedLinear`` weights), trains, and exports its part back in the same layout."""

import json

import pytest
import torch

from tests.dist_utils import run_distributed

pytestmark = pytest.mark.dist


def _config(tmp, source, fmt):
    from d9d_b200.loop.config import TrainerConfig
    from d9d_b200.recipes import Qwen3MoEModelProviderConfig

    model = {"model": {"layer": {"hidden_size": 32, "intermediate_size": 16, "num_experts": 4, "experts_top_k": 2, "num_attention_heads": 4,
                                 "num_key_value_heads": 2, "rms_norm_eps": 1e-6, "head_dim": 8},
                       "num_hidden_layers": 2, "rope_base": 10000.0, "max_position_ids": 64,
                       "split_vocab_size": {"text": 96}, "split_vocab_order": ["text"]}}
    provider = Qwen3MoEModelProviderConfig.model_validate({"model": model, "dtype": "float32", "checkpoint_format": "huggingface",
                                                           "experts_format": fmt})
    trainer = TrainerConfig.model_validate({
        "run": {"name": "hf", "description": None}, "batching": {"global_batch_size": 8, "microbatch_size": 2},
        "data_loading": {"num_workers": 0, "pin_memory": False, "persistent_workers": False},
        "logging": {"period_steps": 1, "tracker": {"provider": "jsonl", "directory": str(tmp / "logs")}},
        "pipelining": {"schedule": {"schedule": "1f1b", "num_stages_per_rank": 1, "zero_bubble": False}},
        "model_stage_factory": {"source_checkpoint": str(source), "checkpoint_only_trainable_parameters": False},
        "determinism": {"base_seed": 0}, "gc": {"period_steps": "disable"},
        "checkpointing": {"save_dir": str(tmp / "ckpt"), "period_steps": "disable", "num_to_keep": None},
        "gradient_clipping": {"max_norm": 1.0, "log_total_steps": 1}, "profiling": None,
        "gradient_manager": {"grad_dtype": None, "bucket_size_mb": 1}})
    return provider, trainer


def _train(tmp, source, fmt, mesh_kwargs):
    from d9d_b200.core.dist_context import DeviceMeshParameters
    from d9d_b200.loop.auto import AutoLRSchedulerProvider, AutoOptimizerProvider
    from d9d_b200.loop.auto.auto_lr_scheduler import PiecewiseConfig
    from d9d_b200.loop.auto.auto_optimizer import AdamWOptimizerConfig
    from d9d_b200.loop.run import TrainingConfigurator
    from d9d_b200.recipes import CausalLMTask, Qwen3MoEModelProvider, SyntheticDataConfig, SyntheticDataProvider

    provider, trainer_cfg = _config(tmp, source, fmt)
    schedule = PiecewiseConfig.model_validate({"name": "piecewise", "scheduler": {"initial_multiplier": 1.0, "phases": [
        {"mode": "rest", "target_multiplier": 1.0, "curve": {"type": "linear"}}]}})
    trainer = TrainingConfigurator(
        mesh=DeviceMeshParameters(**mesh_kwargs), parameters=trainer_cfg, task_provider=lambda ctx: CausalLMTask(ctx.dist_context),
        model_provider=Qwen3MoEModelProvider(provider),
        data_provider=SyntheticDataProvider(SyntheticDataConfig(num_samples=24, seq_len=16, vocab_size=96, seed=1, learnable=True)),
        optimizer_provider=AutoOptimizerProvider(AdamWOptimizerConfig(lr=2e-3, weight_decay=0.0)),
        lr_scheduler_provider=AutoLRSchedulerProvider(schedule)).configure()
    trainer.train()
    trainer.export(tmp / "export", load_checkpoint=False)


def _worker(rank, world, jobs, mesh_kwargs):
    from pathlib import Path

    for tmp, source, fmt in jobs:  # both expert layouts over one set of processes
        _train(Path(tmp), Path(source), fmt, mesh_kwargs)


def _losses(tmp):
    records = [json.loads(line) for line in next((tmp / "logs").glob("*.jsonl")).read_text().splitlines()]
    return {r["step"]: r["value"] for r in records if r.get("name") == "loss"}


def _write_hf_checkpoint(root, fmt):
    import transformers
    from safetensors.torch import save_file

    cfg = transformers.Qwen3MoeConfig(vocab_size=96, hidden_size=32, intermediate_size=48, moe_intermediate_size=16, num_hidden_layers=2,
                                      num_attention_heads=4, num_key_value_heads=2, head_dim=8, rms_norm_eps=1e-6, rope_theta=10000.0,
                                      max_position_embeddings=64, tie_word_embeddings=False, attention_bias=False, num_experts=4,
                                      num_experts_per_tok=2, norm_topk_prob=True, decoder_sparse_step=1, mlp_only_layers=[])
    torch.manual_seed(0)
    state = {k: v.contiguous() for k, v in transformers.Qwen3MoeForCausalLM(cfg).state_dict().items()}
    if fmt == "module_list":  # rewrite the fused expert tensors the way transformers 4 stored them
        for key in [k for k in state if k.endswith("experts.gate_up_proj")]:
            prefix, gate_up, down = key[: -len("gate_up_proj")], state.pop(key), state.pop(key[: -len("gate_up_proj")] + "down_proj")
            for e in range(gate_up.shape[0]):
                gate, up = gate_up[e].chunk(2, dim=0)
                state[f"{prefix}{e}.gate_proj.weight"], state[f"{prefix}{e}.up_proj.weight"] = gate.contiguous(), up.contiguous()
                state[f"{prefix}{e}.down_proj.weight"] = down[e].contiguous()
    root.mkdir(parents=True)
    save_file(state, str(root / "model-00001-of-00001.safetensors"))
    (root / "model.safetensors.index.json").write_text(json.dumps({
        "metadata": {"total_size": sum(v.numel() * v.element_size() for v in state.values())},
        "weight_map": dict.fromkeys(state, "model-00001-of-00001.safetensors")}))
    return state


def test_huggingface_checkpoint_through_a_sharded_pipelined_job(tmp_path):
    pytest.importorskip("transformers")
    from safetensors.torch import load_file

    formats = ("fused", "module_list")
    states = {fmt: _write_hf_checkpoint(tmp_path / fmt / "hf", fmt) for fmt in formats}
    for fmt in formats:
        _train(tmp_path / fmt / "single", tmp_path / fmt / "hf", fmt, {})
    mesh = {"pipeline_parallel": 2, "context_parallel_shard": 2, "expert_parallel": 2}  # all ranks read the same samples
    run_distributed(_worker, 4, [(str(tmp_path / fmt / "dist"), str(tmp_path / fmt / "hf"), fmt) for fmt in formats], mesh)

    def exported(path):
        index = json.loads((path / "model.safetensors.index.json").read_text())
        tensors = {}
        for file in set(index["weight_map"].values()):
            tensors.update(load_file(str(path / file)))
        return tensors

    for fmt in formats:
        ref, got = _losses(tmp_path / fmt / "single"), _losses(tmp_path / fmt / "dist")
        assert sorted(ref) == sorted(got) == [0, 1, 2], fmt
        for step in ref:  # same weights in, same trajectory
            assert abs(ref[step] - got[step]) < 2e-3 * max(1.0, abs(ref[step])), (fmt, step, ref[step], got[step])
        single, dist = exported(tmp_path / fmt / "single" / "export"), exported(tmp_path / fmt / "dist" / "export")
        assert single.keys() == dist.keys() == states[fmt].keys()  # the export is again a complete HuggingFace checkpoint
        for name in single:
            torch.testing.assert_close(dist[name], single[name], rtol=2e-3, atol=2e-4, msg=lambda m, name=name: f"{fmt} {name}: {m}")  # noqa: B023
