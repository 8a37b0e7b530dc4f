"""Context-parallel attention on gloo: ring and Ulysses exchanges against single-process attention over the full
sequence (outputs and gradients of q, k, v), for both token layouts, causal and full attention, GQA head ratios."""

import pytest
import torch

from tests.dist_utils import run_distributed

pytestmark = pytest.mark.dist


def _full_reference(q, k, v, causal):
    from d9d_b200.kernel.flash_attn import attention_reference

    q, k, v = (t.clone().requires_grad_() for t in (q, k, v))
    out, _ = attention_reference(q, k, v, None, causal)
    return q, k, v, out


def _worker(rank, world, mode, heads, kv_heads):
    """One process group, every (layout, causal) combination (spawning processes dominates the cost of these tests)."""
    import torch.distributed as dist

    dist.init_process_group("gloo")
    import d9d_b200.kernel.context_parallel.ring as ring

    ring._QUERY_CHUNK = 5  # noqa: SLF001  several (ragged) query chunks per block even at these tiny sequence lengths
    for layout_name in ("zigzag", "contiguous"):
        for causal in (True, False):
            try:
                _check(rank, world, mode, layout_name, causal, heads, kv_heads)
            except AssertionError as exc:
                raise AssertionError(f"[{mode} layout={layout_name} causal={causal}] {exc}") from exc


def _check(rank, world, mode, layout_name, causal, heads, kv_heads):
    import torch.distributed as dist

    from d9d_b200.kernel.context_parallel import (ContextParallelLayout, gather_sequence, local_sequence_indices, ring_attention,
                                                  shard_sequence, ulysses_attention)
    from d9d_b200.kernel.flash_attn import attention_reference

    group = dist.group.WORLD
    layout = ContextParallelLayout(layout_name)
    torch.manual_seed(0)  # identical full tensors on every rank
    batch, seq, dim = 2, 8 * world, 8
    q = torch.randn(batch, seq, heads, dim)
    k = torch.randn(batch, seq, kv_heads, dim)
    v = torch.randn(batch, seq, kv_heads, dim)
    weight = torch.randn(batch, seq, heads, dim)  # a non-trivial upstream gradient
    q_ref, k_ref, v_ref, out_ref = _full_reference(q, k, v, causal)
    (out_ref * weight).sum().backward()

    local = [shard_sequence(t, 1, world, rank, layout).clone().requires_grad_() for t in (q, k, v)]
    positions = torch.stack([local_sequence_indices(seq, world, r, layout) for r in range(world)])
    if mode in ("ring", "ring_plan"):
        # ring_plan: blocks as flash-attention calls on runs of consecutive positions (the CUDA path, here on the fp32 oracle)
        out = ring_attention(*local, group, positions, causal=causal, mask_cache={}, block_impl="plan" if mode == "ring_plan" else "mask")
    else:
        out = ulysses_attention(*local, group, lambda a, b, c: attention_reference(a, b, c, None, causal)[0],
                                positions=positions.reshape(-1))
    assert out.shape == local[0].shape
    torch.testing.assert_close(out, shard_sequence(out_ref.detach(), 1, world, rank, layout), atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(gather_sequence(out.detach(), 1, group, layout), out_ref.detach(), atol=1e-5, rtol=1e-5)

    (out * shard_sequence(weight, 1, world, rank, layout)).sum().backward()
    for mine, full in zip(local, (q_ref, k_ref, v_ref)):
        torch.testing.assert_close(mine.grad, shard_sequence(full.grad, 1, world, rank, layout), atol=1e-5, rtol=1e-4)


@pytest.mark.parametrize("world,heads,kv_heads", [(2, 4, 2), (4, 4, 1), (3, 3, 3)])
def test_ring_attention_matches_full_attention(world, heads, kv_heads):
    run_distributed(_worker, world, "ring", heads, kv_heads)


@pytest.mark.parametrize("world,heads,kv_heads", [(2, 4, 2), (4, 4, 1)])
def test_ring_attention_plan_blocks_match_full_attention(world, heads, kv_heads):
    run_distributed(_worker, world, "ring_plan", heads, kv_heads)


def test_block_plan_of_the_two_layouts():
    from d9d_b200.kernel.context_parallel import ContextParallelLayout, local_sequence_indices
    from d9d_b200.kernel.context_parallel.ring import _CAUSAL, _FULL, _BlockPlan

    world, seq = 4, 64
    for name in ("zigzag", "contiguous"):
        pos = torch.stack([local_sequence_indices(seq, world, r, ContextParallelLayout(name)) for r in range(world)])
        for rank in range(world):
            plan = _BlockPlan(pos, rank, True)
            for src in range(world):
                calls = plan.calls(src)
                assert calls is not None  # both layouts decompose into full / aligned-causal calls
                visible = sum(1 for _ in calls)
                if name == "contiguous":
                    assert calls == ([(None, None, _CAUSAL)] if src == rank else [(None, None, _FULL)] if src < rank else [])
                elif src == rank:
                    # two runs: (early, early) causal, (late, early) full, (late, late) causal; the last rank's two chunks are adjacent
                    want = [_CAUSAL] if rank == world - 1 else [_FULL, _CAUSAL, _CAUSAL]
                    assert sorted(m for _, _, m in calls) == want and visible == len(want)
                else:
                    # one unmasked call: all my queries x the early run of an earlier rank, or my late run x the whole later block
                    assert len(calls) == 1 and calls[0][2] == _FULL and (calls[0][0] is None) != (calls[0][1] is None) or rank == world - 1 or src == world - 1
    # an irregular layout (interleaved tokens) falls back to the mask path
    weird = torch.stack([torch.arange(r, 16, 2) for r in range(2)])
    assert _BlockPlan(weird, 0, True).calls(1) is None or all(m in (_FULL, _CAUSAL) for _, _, m in _BlockPlan(weird, 0, True).calls(1))


@pytest.mark.parametrize("world,heads,kv_heads", [(2, 4, 2), (4, 8, 4)])
def test_ulysses_attention_matches_full_attention(world, heads, kv_heads):
    run_distributed(_worker, world, "ulysses", heads, kv_heads)


def test_layout_indices_partition_the_sequence():
    from d9d_b200.kernel.context_parallel import ContextParallelLayout, local_sequence_indices

    for layout in ContextParallelLayout:
        parts = [local_sequence_indices(24, 4, r, layout) for r in range(4)]
        assert sorted(torch.cat(parts).tolist()) == list(range(24))
    zig = local_sequence_indices(16, 2, 0, ContextParallelLayout.zigzag).tolist()
    assert zig == [0, 1, 2, 3, 12, 13, 14, 15]  # first and last chunk: equal causal work on every rank
    with pytest.raises(ValueError):
        local_sequence_indices(10, 4, 0, ContextParallelLayout.zigzag)
