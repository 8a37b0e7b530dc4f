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


# ---------------------------------------------------------------------------------------------------------------------
# The optimizer's Python paths (constructor, step, checkpoint) on CPU with the NVLink pieces replaced by stand-ins
# ---------------------------------------------------------------------------------------------------------------------
class _FakeGroup:
    def size(self):
        return 1

    def rank(self):
        return 0


class _FakeArena:
    def __init__(self, numel, dtype, device, group):
        self.buffer = torch.zeros(numel, dtype=dtype, device=device)
        self.peer_ptrs_dev = 0
        self.multicast_ptr = 0

    def barrier(self):
        pass


class _FakeOps:
    """Single-replica semantics of the two kernels (no stochastic rounding): enough to run ``step()`` end to end."""

    def nvl_reduce_shard_(self, grad, peer_ptrs, mc, begin, end, world, rank, sumsq):
        sumsq += grad[begin:end].double().square().sum().float()

    def nvl_adamw_shard_(self, param, grad, m, v, peer_ptrs, mc, begin, end, world, rank, lr, b1, b2, eps, wd, bc1, bc2, seed, scale):
        g = grad[begin:end] * (scale if scale is not None else 1.0)
        mf = b1 * m.float() + (1 - b1) * g
        vf = b2 * v.float() + (1 - b2) * g * g
        p = param[begin:end].float() * (1 - lr * wd) - lr * (mf / bc1) / ((vf / bc2).sqrt() + eps)
        param[begin:end] = p.to(param.dtype)
        m.copy_(mf.to(m.dtype))
        v.copy_(vf.to(v.dtype))


def _patched_optimizer(monkeypatch, params_or_groups, **kw):
    import torch.distributed as dist

    import d9d_b200.optim.nvlink.sharded_adamw as mod

    monkeypatch.setattr(mod, "SymmetricArena", _FakeArena)
    monkeypatch.setattr(mod, "native_ops", lambda: _FakeOps())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(dist, "all_reduce", lambda t, group=None, **k: None)
    return mod.NvlinkShardedAdamW(params_or_groups, _FakeGroup(), lr=1e-2, state_dtype=torch.float32, max_norm=1.0, chunk_numel=2048, **kw)


def _toy_params():
    torch.manual_seed(0)
    return [torch.nn.Parameter(torch.randn(40, 64).bfloat16()), torch.nn.Parameter(torch.randn(64).bfloat16()),
            torch.nn.Parameter(torch.randn(3000).bfloat16())]


def test_single_group_layout_is_the_dense_one_and_a_step_runs(monkeypatch):
    from d9d_b200.optim.nvlink.sharded_adamw import plan_arena_layout

    sizes = [40 * 64, 64, 3000]
    layout = plan_arena_layout([sizes], world=4, chunk_numel=1 << 24)
    dense = sum((n + 7) // 8 * 8 for n in sizes)
    chunk = max(1024, min(1 << 24, -(-(-(-dense // 4)) // 1024) * 1024))
    assert layout.offsets == [0, 2560, 2624] and layout.chunk == chunk and layout.total == -(-dense // (chunk * 4)) * chunk * 4
    assert layout.row_group == [0] * layout.rows and layout.real_numel == 2624 + 3000

    params = _toy_params()
    before = [p.detach().float().clone() for p in params]
    opt = _patched_optimizer(monkeypatch, params)
    assert all(p.grad is not None and p.grad.dtype == torch.float32 for p in params)  # gradients alias the fp32 arena
    for p in params:
        p.grad.copy_(torch.ones_like(p.grad))
    opt.step()
    assert float(opt.last_grad_norm) > 0 and all(float(p.grad.abs().sum()) == 0 for p in params)  # zeroed inside step()
    assert all((p.detach().float() - b).abs().max() > 0 for p, b in zip(params, before))  # every parameter moved
    state = opt.state_dict()
    fresh = _patched_optimizer(monkeypatch, _toy_params())
    fresh.load_state_dict(state)
    torch.testing.assert_close(fresh.exp_avg, opt.exp_avg)
    assert fresh._step_count == 1  # noqa: SLF001


def test_parameter_groups_get_their_own_rows_and_hyper_parameters(monkeypatch):
    params = _toy_params()
    groups = [{"params": [params[0], params[2]]}, {"params": [params[1]], "weight_decay": 0.0, "lr": 0.0}]
    opt = _patched_optimizer(monkeypatch, groups)
    row = opt._chunk * 1  # noqa: SLF001  world = 1
    assert opt._row_group[0] == 0 and opt._row_group[-1] == 1  # noqa: SLF001
    # the second group starts on a row boundary and nothing of the first group shares its chunks
    (lo0, hi0), (lo1, hi1), (lo2, hi2) = opt._param_ranges  # noqa: SLF001  order: group 0 (p0, p2), then group 1 (p1)
    assert lo2 % row == 0 and hi1 <= lo2
    frozen = params[1].detach().float().clone()
    for p in params:
        p.grad.copy_(torch.ones_like(p.grad))
    opt.step()
    torch.testing.assert_close(params[1].detach().float(), frozen)  # lr = 0 in its group: untouched
    assert (params[0].detach().float() - _toy_params()[0].float()).abs().max() > 0
