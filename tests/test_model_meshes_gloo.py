"""Local-vs-parallel self-consistency of whole models over several meshes (CPU/gloo, 4 ranks).

For every mesh the model is parallelised with its family plan, driven through the loop's gradient machinery for one
step, and the *global* gradient of every parameter (DTensors gathered) must equal the single-process gradient over the
whole global batch.  Mirrors the reference's ``modules/model/sequence/*/test_distributed.py`` (8 GPUs there).
"""

import pytest
import torch

from tests.dist_utils import run_distributed

pytestmark = pytest.mark.dist

MESHES = {
    "dpr4": dict(data_parallel_replicate=4),
    "dps4": dict(data_parallel_shard=4),
    "dpr2_dps2": dict(data_parallel_replicate=2, data_parallel_shard=2),
    "dpr4_ep2": dict(data_parallel_replicate=4, expert_parallel=2),
    "dpr2_dps2_ep4": dict(data_parallel_replicate=2, data_parallel_shard=2, expert_parallel=4),
    # context parallelism: ranks of one cp group read the same samples and keep S / cp tokens of every sequence
    "cps4": dict(context_parallel_shard=4),
    "cpr2_cps2": dict(context_parallel_replicate=2, context_parallel_shard=2),
    "dpr2_cps2": dict(data_parallel_replicate=2, context_parallel_shard=2),
    "dps2_cps2_ep2": dict(data_parallel_shard=2, context_parallel_shard=2, expert_parallel=2),
    # tensor parallelism (with sequence parallelism): heads and MLP columns split over tp, S / tp tokens between blocks
    "tp2": dict(tensor_parallel=2),
    "dpr2_tp2": dict(data_parallel_replicate=2, tensor_parallel=2),
    "cpr2_tp2": dict(context_parallel_replicate=2, tensor_parallel=2),
    "dpr2_tp2_ep2": dict(data_parallel_replicate=2, tensor_parallel=2, expert_parallel=2),
    "dps2_tp2": dict(data_parallel_shard=2, tensor_parallel=2),  # FSDP over tensor-parallel parameters (strided shards)
}


def _build(moe: bool):
    from tests.helpers_train import dense_params, moe_params

    from d9d_b200.module.block.hidden_states_aggregator import HiddenStatesAggregationMode
    from d9d_b200.pipelining.api import PipelineStageInfo

    torch.manual_seed(5)
    if moe == "deepseek":
        from d9d_b200.module.model.deepseek_v2 import (DeepseekV2ForCausalLM as Cls, DeepseekV2ForCausalLMParameters,
                                                       DeepseekV2LayerParameters, DeepseekV2Parameters)
        from d9d_b200.module.model.deepseek_v3 import deepseek_v3_router

        params = DeepseekV2ForCausalLMParameters(model=DeepseekV2Parameters(
            layer=DeepseekV2LayerParameters(hidden_size=32, rms_norm_eps=1e-6, num_attention_heads=4, qk_nope_head_dim=8,
                                            qk_rope_head_dim=4, v_head_dim=8, kv_lora_rank=16, q_lora_rank=12, intermediate_size=48,
                                            first_k_dense_replace=1, moe_intermediate_size=16, num_experts=4, experts_top_k=2,
                                            num_shared_experts=1, router_renormalize_probabilities=True,
                                            router=deepseek_v3_router(n_group=2, topk_group=1, routed_scaling_factor=2.5)),
            num_hidden_layers=2, rope_base=10000, max_position_ids=64, split_vocab_size={"regular": 100, "special": 28},
            split_vocab_order=["regular", "special"]))
    elif moe == "qwen3_5_moe":
        from d9d_b200.module.model.qwen3_5_moe import (Qwen3_5MoEForCausalLM as Cls, Qwen3_5MoEForCausalLMParameters,
                                                       Qwen3_5MoELayerParameters, Qwen3_5MoEParameters)

        params = Qwen3_5MoEForCausalLMParameters(model=Qwen3_5MoEParameters(
            layer=Qwen3_5MoELayerParameters(hidden_size=32, rms_norm_eps=1e-6, num_attention_heads=4, num_key_value_heads=2, head_dim=16,
                                            linear_num_key_heads=2, linear_num_value_heads=4, linear_key_head_dim=8,
                                            linear_value_head_dim=8, full_attention_interval=2, moe_intermediate_size=16,
                                            shared_expert_intermediate_size=24, num_experts=4, experts_top_k=2),
            num_hidden_layers=2, rope_base=10000, max_position_ids=64, split_vocab_size={"regular": 100, "special": 28},
            split_vocab_order=["regular", "special"]))
    elif moe:
        from d9d_b200.module.model.qwen3_moe import Qwen3MoEForCausalLM as Cls

        params = moe_params()
    else:
        from d9d_b200.module.model.qwen3_dense import Qwen3DenseForCausalLM as Cls

        params = dense_params()
    model = Cls(params, PipelineStageInfo(0, 1), HiddenStatesAggregationMode.no, False)
    model.reset_parameters()
    return model


def _batch(i):
    g = torch.Generator().manual_seed(900 + i)
    ids = torch.randint(0, 128, (2, 16), generator=g)
    labels = torch.randint(0, 128, (2, 16), generator=g)
    return ids, labels, torch.arange(16)[None].expand(2, -1)


def _worker(rank, world_size, mesh_names, moe):
    """Several meshes over one set of processes (spawning processes dominates the cost of these tests)."""
    for mesh_name in mesh_names:
        _check_mesh(rank, world_size, mesh_name, moe)


def _check_mesh(rank, world_size, mesh_name, moe):
    import torch.distributed as dist
    from torch.distributed.tensor import DTensor

    from d9d_b200.core.dist_context import BATCH_DOMAIN, DeviceMeshParameters
    from d9d_b200.internals.grad_sync import GradientSynchronizer
    from d9d_b200.pipelining.api import PipelineStageInfo

    ctx = DeviceMeshParameters(**MESHES[mesh_name]).build()
    model = _build(moe)
    if moe == "qwen3_5_moe":
        from d9d_b200.module.parallelism.model.qwen3_5_moe import parallelize_qwen3_5_moe_for_causal_lm as plan
    elif moe == "deepseek":
        from d9d_b200.module.parallelism.model.deepseek_v2 import parallelize_deepseek_v2_for_causal_lm as plan
    elif moe:
        from d9d_b200.module.parallelism.model.qwen3_moe import parallelize_qwen3_moe_for_causal_lm as plan
    else:
        from d9d_b200.module.parallelism.model.qwen3_dense import parallelize_qwen3_dense_for_causal_lm as plan
    plan(ctx, model, PipelineStageInfo(0, 1))

    params = list(model.parameters())
    sync = GradientSynchronizer([params], bucket_size_mb=1, require_accumulations=1)
    sync.bind()
    from d9d_b200.dataset import shard_batch_along_sequence

    batch_mesh = ctx.mesh_for(BATCH_DOMAIN)
    dp_rank = batch_mesh["dp"].get_local_rank()
    ids, labels, pos = shard_batch_along_sequence(_batch(dp_rank), ctx)  # identity without context / tensor parallelism
    assert ids.shape[1] == 16 // (batch_mesh["cp"].size() * batch_mesh["tp"].size())
    model(input_ids=ids, position_ids=pos, labels=labels)["logps"].sum().backward()
    sync.wait()

    ref = _build(moe)
    for b in range(batch_mesh["dp"].size()):  # one batch per data-parallel replica
        ids, labels, pos = _batch(b)
        ref(input_ids=ids, position_ids=pos, labels=labels)["logps"].sum().backward()

    for (name, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        assert p.grad is not None, name
        got = p.grad.full_tensor() if isinstance(p.grad, DTensor) else p.grad
        torch.testing.assert_close(got, q.grad, rtol=2e-4, atol=2e-5, msg=lambda m: f"[{mesh_name}] {name}: {m}")  # noqa: B023
    dist.barrier()


@pytest.mark.parametrize("mesh_names", [("dpr4", "dps4", "dpr2_dps2"), ("cps4", "cpr2_cps2", "dpr2_cps2"), ("dpr2_tp2", "cpr2_tp2", "dps2_tp2")],
                         ids="+".join)
def test_dense_model_matches_single_process(mesh_names):
    run_distributed(_worker, 4, mesh_names, False)


def test_dense_model_matches_single_process_on_two_tensor_parallel_ranks():
    run_distributed(_worker, 2, ("tp2",), False)


@pytest.mark.parametrize("mesh_names", [("dpr4", "dps4", "dpr4_ep2"), ("dpr2_dps2_ep4", "dps2_cps2_ep2", "dpr2_tp2_ep2")], ids="+".join)
def test_moe_model_matches_single_process(mesh_names):
    run_distributed(_worker, 4, mesh_names, True)


def test_deepseek_v2_model_matches_single_process():
    """Latent attention under context parallelism, dense first layer + MoE layers (DeepSeek-V3 routing: sigmoid scores, selection
    bias, group-limited top-k) with a shared expert under FSDP x EP."""
    run_distributed(_worker, 4, ("dps2_cps2_ep2",), "deepseek")


def test_qwen3_5_moe_model_matches_single_process():
    """Gated DeltaNet / gated attention hybrid with MoE + gated shared expert under HSDP x EP and FSDP."""
    run_distributed(_worker, 4, ("dpr2_dps2_ep4", "dps4"), "qwen3_5_moe")
