"""LoRA adapters on attention projections and on the grouped experts of a MoE model under HSDP x expert parallelism (gloo):
the adapter weights of the experts are sharded along the expert dim like their base weights, and the gradients of all
adapters equal the single-process ones."""

import pytest
import torch

from tests.dist_utils import run_distributed

pytestmark = pytest.mark.dist


def _build():
    from tests.test_model_meshes_gloo import _build as build
    from d9d_b200.peft import inject_peft_and_freeze
    from d9d_b200.peft.all import peft_method_from_config
    from d9d_b200.peft.lora.config import LoRAConfig
    m = build(True)
    method = peft_method_from_config(LoRAConfig.model_validate({"kind": "lora", "module_name_pattern": r".*(grouped_experts\.(gate_proj|up_proj|down_proj)|self_attn\.q_proj)", "params": {"r": 2, "alpha": 4, "dropout": 0.0}}))
    torch.manual_seed(11)
    inject_peft_and_freeze(method, m)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "lora_B" in n:
                p.normal_(0, 0.1)
    return m

def _w(rank, world):
    from torch.distributed.tensor import DTensor
    from d9d_b200.core.dist_context import BATCH_DOMAIN, DeviceMeshParameters
    from d9d_b200.internals.grad_sync import GradientSynchronizer
    from d9d_b200.module.parallelism.model.qwen3_moe import parallelize_qwen3_moe_for_causal_lm as plan
    from d9d_b200.pipelining.api import PipelineStageInfo
    from tests.test_model_meshes_gloo import _batch
    ctx = DeviceMeshParameters(data_parallel_replicate=2, data_parallel_shard=2, expert_parallel=2).build()
    model = _build()
    plan(ctx, model, PipelineStageInfo(0, 1))
    params = [p for p in model.parameters() if p.requires_grad]
    sync = GradientSynchronizer([params], bucket_size_mb=1, require_accumulations=1)
    sync.bind()
    dp = ctx.mesh_for(BATCH_DOMAIN)["dp"]
    ids, labels, pos = _batch(dp.get_local_rank())
    model(input_ids=ids, position_ids=pos, labels=labels)["logps"].sum().backward()
    sync.wait()
    ref = _build()
    for b in range(dp.size()):
        ids, labels, pos = _batch(b)
        ref(input_ids=ids, position_ids=pos, labels=labels)["logps"].sum().backward()
    n_checked = 0
    for (name, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        if not q.requires_grad:
            continue
        got = p.grad.full_tensor() if isinstance(p.grad, DTensor) else p.grad
        torch.testing.assert_close(got, q.grad, rtol=2e-4, atol=2e-5, msg=lambda m, name=name: f"{name}: {m}")
        n_checked += 1
    assert n_checked > 0

def test_lora_on_experts_under_expert_parallelism():
    run_distributed(_w, 4)
