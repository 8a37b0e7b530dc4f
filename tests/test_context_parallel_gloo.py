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
    if mode == "ring":
        out = ring_attention(*local, group, positions, causal=causal, mask_cache={})
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
