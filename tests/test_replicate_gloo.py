"""BASELINE config #1: tiny 2-layer dense LM, parallelize_replicate on CPU/gloo world_size=2 (plumbing, no GPU)."""

import pytest
import torch

from tests.dist_utils import run_distributed

pytestmark = pytest.mark.dist


def _build():
    from d9d_b200.module.block.hidden_states_aggregator import HiddenStatesAggregationMode
    from d9d_b200.module.model.qwen3_dense import (Qwen3DenseForCausalLM, Qwen3DenseForCausalLMParameters,
                                                   Qwen3DenseLayerParameters, Qwen3DenseParameters)
    from d9d_b200.pipelining.api import PipelineStageInfo

    p = Qwen3DenseForCausalLMParameters(model=Qwen3DenseParameters(
        layer=Qwen3DenseLayerParameters(hidden_size=32, intermediate_size=64, num_attention_heads=4, num_key_value_heads=2,
                                        rms_norm_eps=1e-6, head_dim=8),
        num_hidden_layers=2, rope_base=10000, max_position_ids=64,
        split_vocab_size={"regular": 50, "special": 14}, split_vocab_order=["regular", "special"]))
    torch.manual_seed(7)
    m = Qwen3DenseForCausalLM(p, PipelineStageInfo(0, 1), HiddenStatesAggregationMode.no, False)
    m.reset_parameters()
    return m


def _batch(rank):
    g = torch.Generator().manual_seed(100 + rank)
    ids = torch.randint(0, 64, (2, 16), generator=g)
    labels = torch.randint(0, 64, (2, 16), generator=g)
    pos = torch.arange(16)[None].expand(2, -1)
    return ids, labels, pos


def _worker(rank, world_size):
    from torch.distributed.tensor import DTensor, Replicate

    from d9d_b200.core.dist_context import DENSE_DOMAIN, DeviceMeshParameters
    from d9d_b200.internals.grad_norm import clip_grad_norm_distributed_, group_parameters_for_norm
    from d9d_b200.internals.grad_sync import GradientSynchronizer
    from d9d_b200.module.parallelism.api import parallelize_replicate

    ctx = DeviceMeshParameters(data_parallel_replicate=world_size).build()
    assert ctx.mesh_for(DENSE_DOMAIN)["dp_replicate"].size() == world_size

    model = _build()
    parallelize_replicate(model, ctx.mesh_for(DENSE_DOMAIN)["dp_replicate"])
    params = list(model.parameters())
    assert all(isinstance(p.data, DTensor) and p.placements == (Replicate(),) for p in params)

    sync = GradientSynchronizer([params], bucket_size_mb=1, require_accumulations=2)
    sync.bind()
    assert all(p.grad is not None and float(p.grad.to_local().abs().sum()) == 0 for p in params)  # zero-init aliasing

    # two accumulation rounds; after the first the buckets must NOT be ready
    for round_idx in range(2):
        ids, labels, pos = _batch(rank * 2 + round_idx)
        out = model(input_ids=ids, position_ids=pos, labels=labels)
        out["logps"].sum().backward()
        if round_idx == 0:
            with pytest.raises(ValueError):
                sync.wait()
    sync.wait()

    # single-process reference over all 4 batches (gradients are SUMmed)
    ref = _build()
    for b in range(4):
        ids, labels, pos = _batch(b)
        ref(input_ids=ids, position_ids=pos, labels=labels)["logps"].sum().backward()
    for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        torch.testing.assert_close(p.grad.to_local(), q.grad, rtol=1e-4, atol=1e-5, msg=lambda m: f"{n}: {m}")  # noqa: B023

    # distributed norm == local norm of the summed grads; clipping scales in place
    groups = group_parameters_for_norm(params)
    total = clip_grad_norm_distributed_(groups, max_norm=None, norm_type=2.0, pp_mesh=None)
    ref_norm = torch.nn.utils.get_total_norm([q.grad for q in ref.parameters()])
    torch.testing.assert_close(total, ref_norm, rtol=1e-4, atol=1e-5)
    clip_grad_norm_distributed_(groups, max_norm=0.5 * float(ref_norm), norm_type=2.0, pp_mesh=None)
    new_norm = torch.nn.utils.get_total_norm([p.grad.to_local() for p in params])
    torch.testing.assert_close(new_norm, 0.5 * ref_norm, rtol=1e-3, atol=1e-5)

    sync.zero_grad()
    assert all(float(p.grad.to_local().abs().sum()) == 0 for p in params)
    sync.unbind()
    assert all(p.grad is None for p in params)
    # state dict keeps DTensors (what DCP / export rely on)
    assert all(isinstance(v, DTensor) for k, v in model.state_dict().items() if "weight" in k)


def test_replicate_grad_sync_gloo():
    run_distributed(_worker, 2)
