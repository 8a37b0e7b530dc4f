"""Checkpoint layout of ``NvlinkShardedAdamW`` (pure host logic: the optimizer itself needs NVLink peer memory): moments
are keyed by global chunk id, so a checkpoint written by W replicas restores on W' replicas."""

import torch

from d9d_b200.optim.nvlink.sharded_adamw import chunk_state_views, load_chunk_state, owned_real_chunks


def _layout(real_numel: int, chunk: int, world: int):
    row = chunk * world
    total = -(-real_numel // row) * row
    return total // row  # chunks (rows) per rank


def _global_moments(real_numel, chunk, world, fill):
    """Per-rank moment buffers holding ``fill(global element index)`` on real elements."""
    rows = _layout(real_numel, chunk, world)
    ranks = []
    for rank in range(world):
        m = torch.zeros(rows * chunk)
        for c, slot in owned_real_chunks(chunk, world, rank, real_numel, rows):
            idx = torch.arange(c * chunk, (c + 1) * chunk)
            m[slot * chunk : (slot + 1) * chunk] = torch.where(idx < real_numel, fill(idx), torch.zeros(()))
        ranks.append(m)
    return ranks


def test_every_real_chunk_has_exactly_one_owner():
    real, chunk = 10_500, 1024
    for world in (1, 2, 3, 8):
        rows = _layout(real, chunk, world)
        owned = sorted(c for r in range(world) for c, _ in owned_real_chunks(chunk, world, r, real, rows))
        assert owned == list(range(-(-real // chunk)))


def test_checkpoint_written_by_two_replicas_restores_on_four_and_one():
    real, chunk = 10_500, 1024
    fill = lambda i: i.float() * 0.5 + 1.0  # noqa: E731
    saved: dict[str, dict[str, torch.Tensor]] = {}
    for rank, m in enumerate(_global_moments(real, chunk, 2, fill)):
        saved.update(chunk_state_views(m, m * 2, chunk, 2, rank, real))  # DCP merges the ranks' disjoint keys
    for world in (4, 1, 3):
        rows = _layout(real, chunk, world)
        want = _global_moments(real, chunk, world, fill)
        for rank in range(world):
            m, v = torch.full((rows * chunk,), -1.0), torch.full((rows * chunk,), -1.0)
            load_chunk_state(saved, m, v, chunk, world, rank, real)
            for _c, slot in owned_real_chunks(chunk, world, rank, real, rows):
                sl = slice(slot * chunk, (slot + 1) * chunk)
                torch.testing.assert_close(m[sl], want[rank][sl])
                torch.testing.assert_close(v[sl], want[rank][sl] * 2)
