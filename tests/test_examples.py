"""Smoke tests of the example scripts: each runs end to end on the CPU with a shrunken copy of its shipped config."""

import importlib.util
import json
import sys
from pathlib import Path

import pytest

EXAMPLES = Path(__file__).resolve().parent.parent / "example"


def _load(path: Path):
    spec = importlib.util.spec_from_file_location(f"example_{path.stem}", path)
    module = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = module  # pydantic resolves forward references through sys.modules
    spec.loader.exec_module(module)
    return module


def test_lora_classification_example(tmp_path, monkeypatch):
    config = json.loads((EXAMPLES / "lora_classification" / "finetune.json").read_text())
    config["data"]["num_samples"] = 256
    config["model"]["model"]["layer"].update(hidden_size=32, intermediate_size=64, num_attention_heads=4, num_key_value_heads=2, head_dim=8)
    config["model"]["model"]["num_hidden_layers"] = 2
    config["trainer"]["batching"] = {"global_batch_size": 32, "microbatch_size": 16}
    config["trainer"]["logging"]["tracker"]["directory"] = str(tmp_path / "logs")
    config["trainer"]["checkpointing"]["save_dir"] = str(tmp_path / "ckpt")
    config["export_to"] = str(tmp_path / "export")
    path = tmp_path / "finetune.json"
    path.write_text(json.dumps(config))
    _load(EXAMPLES / "lora_classification" / "finetune.py").main([str(path)])

    records = [json.loads(line) for line in next((tmp_path / "logs").glob("*.jsonl")).read_text().splitlines()]
    names = {r.get("name") for r in records}
    assert {"loss", "accuracy", "f1_macro"} <= names
    losses = [r["value"] for r in records if r.get("name") == "loss"]
    assert all(v == v for v in losses) and losses[-1] < losses[0]  # finite (the frozen base is initialised) and learning
    index = json.loads((tmp_path / "export" / "model.safetensors.index.json").read_text())
    assert "score.weight" in index["weight_map"] and "model.layers.0.self_attn.q_proj.weight" in index["weight_map"]
    assert not any("lora" in key for key in index["weight_map"])  # adapters were merged before the export


def test_qwen3_moe_pretrain_example(tmp_path, monkeypatch):
    config = json.loads((EXAMPLES / "qwen3_moe" / "pretrain.json").read_text())
    config["data"].update(num_samples=32, seq_len=16, vocab_size=64)
    model = config["model_provider"]["model"]["model"]
    model["layer"].update(hidden_size=32, intermediate_size=16, num_experts=4, experts_top_k=2, num_attention_heads=4, num_key_value_heads=2, head_dim=8)
    model.update(num_hidden_layers=2, max_position_ids=64, split_vocab_size={"regular": 60, "special": 4})
    config["model_provider"]["dtype"] = "float32"
    config["trainer"]["batching"] = {"global_batch_size": 8, "microbatch_size": 4}
    config["trainer"]["data_loading"].update(num_workers=0, pin_memory=False)
    config["trainer"]["logging"]["tracker"] = {"provider": "jsonl", "directory": str(tmp_path / "logs")}
    config["trainer"]["checkpointing"]["save_dir"] = str(tmp_path / "ckpt")
    config["trainer"]["profiling"] = None
    if config["optimizer"]["name"] in ("stochastic_adamw", "nvlink_sharded_adamw"):
        config["optimizer"] = {"name": "adamw", "lr": 1e-3}
    config["export_to"] = str(tmp_path / "export")
    path = tmp_path / "pretrain.json"
    path.write_text(json.dumps(config))
    monkeypatch.setattr(sys, "argv", ["pretrain.py", str(path), "--single"])
    _load(EXAMPLES / "qwen3_moe" / "pretrain.py").main()
    assert (tmp_path / "export" / "model.safetensors.index.json").exists()
    assert any((tmp_path / "logs").glob("*.jsonl"))


def test_qwen3_moe_perplexity_example(tmp_path, monkeypatch, capsys):
    config = json.loads((EXAMPLES / "qwen3_moe" / "calculate_perplexity.json").read_text())
    config["data"].update(num_samples=16, seq_len=16, vocab_size=64)
    model = config["model_provider"]["model"]["model"]
    model["layer"].update(hidden_size=32, intermediate_size=16, num_experts=4, experts_top_k=2, num_attention_heads=4, num_key_value_heads=2, head_dim=8)
    model.update(num_hidden_layers=2, max_position_ids=64, split_vocab_size={"regular": 60, "special": 4})
    config["model_provider"]["dtype"] = "float32"
    config["inference"]["batching"] = {"global_batch_size": 8, "microbatch_size": 4}
    config["inference"]["data_loading"].update(num_workers=0, pin_memory=False)
    config["inference"]["model_stage_factory"]["source_checkpoint"] = None
    config["inference"]["checkpointing"]["save_dir"] = str(tmp_path / "progress")
    config["inference"]["profiling"] = None
    path = tmp_path / "perplexity.json"
    path.write_text(json.dumps(config))
    monkeypatch.setattr(sys, "argv", ["calculate_perplexity.py", str(path), "--single"])
    monkeypatch.syspath_prepend(str(EXAMPLES / "qwen3_moe"))  # the script imports its sibling ``pretrain`` like when run in place
    _load(EXAMPLES / "qwen3_moe" / "calculate_perplexity.py").main()
    out = capsys.readouterr().out
    assert "tokens=256" in out and "perplexity=" in out
