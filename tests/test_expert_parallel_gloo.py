"""Expert parallelism on CPU/gloo: a MoE LM sharded EP=2 (DP=2) must reproduce the single-process gradients."""

import pytest
import torch

from tests.dist_utils import run_distributed

pytestmark = pytest.mark.dist


def _build():
    from d9d_b200.module.block.hidden_states_aggregator import HiddenStatesAggregationMode
    from d9d_b200.module.model.qwen3_moe import (Qwen3MoEForCausalLM, Qwen3MoEForCausalLMParameters,
                                                 Qwen3MoELayerParameters, Qwen3MoEParameters)
    from d9d_b200.pipelining.api import PipelineStageInfo

    p = Qwen3MoEForCausalLMParameters(model=Qwen3MoEParameters(
        layer=Qwen3MoELayerParameters(hidden_size=32, intermediate_size=16, num_experts=8, experts_top_k=2,
                                      num_attention_heads=4, num_key_value_heads=2, rms_norm_eps=1e-6, head_dim=8),
        num_hidden_layers=2, rope_base=10000, max_position_ids=64,
        split_vocab_size={"regular": 50, "special": 14}, split_vocab_order=["regular", "special"]))
    torch.manual_seed(11)
    m = Qwen3MoEForCausalLM(p, PipelineStageInfo(0, 1), HiddenStatesAggregationMode.no, False)
    m.reset_parameters()
    return m


def _batch(idx):
    g = torch.Generator().manual_seed(300 + idx)
    ids = torch.randint(0, 64, (2, 16), generator=g)
    labels = torch.randint(0, 64, (2, 16), generator=g)
    pos = torch.arange(16)[None].expand(2, -1)
    return ids, labels, pos


def _worker(rank, world_size):
    from torch.distributed.tensor import DTensor

    from d9d_b200.core.dist_context import DeviceMeshParameters
    from d9d_b200.internals.grad_sync import GradientSynchronizer
    from d9d_b200.module.block.moe import MoELayer
    from d9d_b200.module.block.moe.communications import AutoExpertParallelCommunicationHandler
    from d9d_b200.module.parallelism.model.qwen3_moe import parallelize_qwen3_moe_for_causal_lm
    from d9d_b200.pipelining.api import PipelineStageInfo

    ctx = DeviceMeshParameters(data_parallel_replicate=world_size, expert_parallel=world_size).build()
    model = _build()
    parallelize_qwen3_moe_for_causal_lm(ctx, model, PipelineStageInfo(0, 1))
    moe_layers = [m for m in model.modules() if isinstance(m, MoELayer)]
    assert moe_layers and all(isinstance(m._communicator, AutoExpertParallelCommunicationHandler) for m in moe_layers)
    w = moe_layers[0].grouped_experts.gate_proj.weight
    assert isinstance(w.data, DTensor) and w.to_local().shape[0] == 8 // world_size

    params = list(model.parameters())
    sync = GradientSynchronizer([params], bucket_size_mb=1, require_accumulations=1)
    sync.bind()
    ids, labels, pos = _batch(rank)
    out = model(input_ids=ids, position_ids=pos, labels=labels)
    out["logps"].sum().backward()
    sync.wait()

    ref = _build()
    ref_out = []
    for b in range(world_size):
        ids, labels, pos = _batch(b)
        o = ref(input_ids=ids, position_ids=pos, labels=labels)["logps"]
        ref_out.append(o)
        o.sum().backward()
    torch.testing.assert_close(out["logps"], ref_out[rank], rtol=1e-4, atol=1e-5)

    local_experts = 8 // world_size
    for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        got = p.grad.to_local() if isinstance(p.grad, DTensor) else p.grad
        want = q.grad
        if got.shape != want.shape:  # expert-sharded weights: this rank owns a contiguous block of experts
            want = want[rank * local_experts : (rank + 1) * local_experts]
        torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5, msg=lambda m: f"{n}: {m}")  # noqa: B023


def test_expert_parallel_matches_single_process():
    run_distributed(_worker, world_size=2)
