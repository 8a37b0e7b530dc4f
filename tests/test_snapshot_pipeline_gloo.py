"""Per-layer hidden-state snapshots and the embedding head carried through a 2-stage pipeline (gloo): the last stage must
see the masked means of *all* layers (those of stage 0 arrive over the stage boundary) and produce the single-process
embeddings; gradients flow back through both stages."""

import pytest
import torch

from tests.dist_utils import run_distributed

pytestmark = pytest.mark.dist


def _params():
    from d9d_b200.module.model.qwen3_dense import Qwen3DenseForEmbeddingParameters, Qwen3DenseLayerParameters, Qwen3DenseParameters

    return Qwen3DenseForEmbeddingParameters(model=Qwen3DenseParameters(
        layer=Qwen3DenseLayerParameters(hidden_size=32, intermediate_size=64, num_attention_heads=4, num_key_value_heads=2,
                                        rms_norm_eps=1e-6, head_dim=8),
        num_hidden_layers=4, rope_base=10000, max_position_ids=64, split_vocab_size={"text": 50}, split_vocab_order=["text"]),
        embedding_dim=16, normalize=True)


def _build(stage):
    from d9d_b200.module.block.hidden_states_aggregator import HiddenStatesAggregationMode
    from d9d_b200.module.model.qwen3_dense import Qwen3DenseForEmbedding

    torch.manual_seed(100 + stage.current_stage)  # every stage initialises its own parameters reproducibly
    model = Qwen3DenseForEmbedding(_params(), stage, HiddenStatesAggregationMode.mean, False)
    model.reset_parameters()
    return model


def _batch():
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, 50, (4, 10), generator=g)
    mask = torch.ones(4, 10, dtype=torch.long)
    mask[1, 6:] = 0
    mask[3, 3:] = 0
    pooling = torch.zeros_like(mask)
    pooling[torch.arange(4), mask.sum(1) - 1] = 1  # last real token of every row
    return {"input_ids": ids}, {"position_ids": torch.arange(10).expand(4, -1), "hidden_states_agg_mask": mask, "pooling_mask": pooling}


def _worker(rank, world):
    from d9d_b200.core.dist_context import DeviceMeshParameters
    from d9d_b200.pipelining.api import PipelineStageInfo
    from d9d_b200.pipelining.factory import PipelineScheduleGPipeConfig, build_schedule

    ctx = DeviceMeshParameters(pipeline_parallel=2).build()
    seen = {}

    def loss_fn(outputs, microbatch):
        seen[microbatch] = {k: v.detach().clone() for k, v in outputs.items() if v is not None}
        return outputs["embeddings"].square().sum() + outputs["hidden_states_snapshot"].square().sum()

    info, modules = build_schedule(ctx, n_microbatches=2, schedule_config=PipelineScheduleGPipeConfig(), model_provider=_build,
                                   callback=loss_fn)
    inputs, kwargs = _batch()
    info.schedule.configure_buffers(inputs, kwargs, None)
    info.schedule.step(inputs, kwargs)

    # single-process reference: both stages chained by hand
    stages = [_build(PipelineStageInfo(s, 2)) for s in range(2)]
    first = stages[0](**inputs, **kwargs)
    last = stages[1](hidden_states=first["hidden_states"], hidden_states_snapshot=first["hidden_states_snapshot"], **kwargs)
    assert last["hidden_states_snapshot"].shape == (5, 4, 32)  # embeddings + 4 layers, [entries, batch, hidden]
    (last["embeddings"].square().sum() + last["hidden_states_snapshot"].square().sum()).backward()

    if info.has_last_stage:
        got_snapshot = torch.cat([seen[m]["hidden_states_snapshot"] for m in range(2)], dim=1)
        got_embeddings = torch.cat([seen[m]["embeddings"] for m in range(2)], dim=0)
        torch.testing.assert_close(got_snapshot, last["hidden_states_snapshot"].detach(), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(got_embeddings, last["embeddings"].detach(), rtol=1e-5, atol=1e-6)
        assert torch.allclose(got_embeddings.norm(dim=-1), torch.ones(4), atol=1e-5)  # normalised projection of the pooled tokens
    mine = stages[modules[0]._stage.current_stage]
    for (name, p), (_, q) in zip(modules[0].named_parameters(), mine.named_parameters()):
        torch.testing.assert_close(p.grad, q.grad, rtol=1e-4, atol=1e-5, msg=lambda m, name=name: f"{name}: {m}")


def test_snapshots_and_embedding_head_through_a_pipeline():
    run_distributed(_worker, 2)
