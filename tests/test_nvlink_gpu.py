"""Multi-GPU tests of the NVLink peer-memory kernels (need >= 2 GPUs; skipped otherwise)."""

import os

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]  # a hung collective must not burn GPU-minutes


def _need_gpus(n):
    if not torch.cuda.is_available() or torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")


def _spawn(fn, world, *args):
    import torch.multiprocessing as mp

    port = 29600 + (os.getpid() % 300)
    mp.spawn(_entry, args=(world, port, fn, args), nprocs=world, join=True)


def _entry(rank, world, port, fn, args):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    try:
        fn(rank, world, *args)
    finally:
        dist.destroy_process_group()


def _sharded_adamw_worker(rank, world, state_dtype, max_norm, overlap=False):
    import torch.distributed as dist

    from d9d_b200.optim.nvlink import NvlinkShardedAdamW

    torch.manual_seed(0)  # identical parameters on every replica
    shapes = [(300, 64), (1000,), (7, 24, 40), (129,)]
    params = [torch.nn.Parameter(torch.randn(s, device="cuda").bfloat16()) for s in shapes]
    ref_p = [p.detach().float().clone() for p in params]
    opt = NvlinkShardedAdamW(params, dist.group.WORLD, lr=1e-2, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1,
                             state_dtype=state_dtype, max_norm=max_norm, seed=5, chunk_numel=2048, overlap_waves=3)
    if overlap:
        opt.set_required_accumulations(1)
    assert all(p.grad is not None and p.grad.dtype == torch.float32 for p in params)
    for p, r in zip(params, ref_p):
        torch.testing.assert_close(p.detach().float(), r)  # moving into the arena kept the values

    ref_m = [torch.zeros_like(r) for r in ref_p]
    ref_v = [torch.zeros_like(r) for r in ref_p]
    for step in range(1, 4):
        all_grads = []
        for r in range(world):
            g = torch.Generator(device="cuda").manual_seed(100 * step + r)
            all_grads.append([torch.randn(s, device="cuda", generator=g) for s in shapes])
        for i in reversed(range(len(params))):  # backward produces gradients back to front
            params[i].grad.add_(all_grads[rank][i])  # what backward would accumulate on this replica
            if overlap:
                opt._on_grad_final(i)  # the post-accumulate hook autograd would fire: finished waves reduce right away
        opt.grad_scale = torch.full((1,), 0.5, device="cuda")
        opt.step()
        # fp32 reference of: sum over replicas -> *0.5 -> clip -> AdamW
        summed = [sum(all_grads[r][i] for r in range(world)) * 0.5 for i in range(len(shapes))]
        if max_norm is not None:
            norm = torch.sqrt(sum((g**2).sum() for g in summed))
            torch.testing.assert_close(opt.last_grad_norm.reshape(()), norm, rtol=1e-4, atol=1e-4)
            coef = torch.clamp(max_norm / (norm + 1e-6), max=1.0)
            summed = [g * coef for g in summed]
        for i, g in enumerate(summed):
            ref_p[i] *= 1 - 1e-2 * 0.1
            ref_m[i] = 0.9 * ref_m[i] + 0.1 * g
            ref_v[i] = 0.95 * ref_v[i] + 0.05 * g * g
            ref_p[i] -= 1e-2 * (ref_m[i] / (1 - 0.9**step)) / ((ref_v[i] / (1 - 0.95**step)).sqrt() + 1e-8)
        for p, r in zip(params, ref_p):
            diff = p.detach().float() - r
            assert diff.abs().max() < 0.04, diff.abs().max()  # one bf16 ulp at |p| ~ 4 (stochastic rounding)
            if r.numel() > 5000:
                assert abs(float(diff.mean())) < 1.5e-3  # ... and unbiased
            assert float(p.grad.abs().max()) == 0.0  # gradients were zeroed
        # every replica holds bit-identical parameters
        flat = torch.cat([p.detach().float().flatten() for p in params])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        for other in gathered:
            assert torch.equal(other, gathered[0])
        # keep the fp32 reference from drifting away from the bf16 trajectory
        ref_p = [p.detach().float().clone() for p in params]


@pytest.mark.parametrize("state_dtype,max_norm,multimem,overlap", [
    (torch.bfloat16, 1.0, "0", True), (torch.bfloat16, None, "1", False), (torch.float32, 1.0, "1", True), (torch.float32, None, "0", False)])
def test_nvlink_sharded_adamw_matches_reference(state_dtype, max_norm, multimem, overlap, monkeypatch):
    _need_gpus(2)
    monkeypatch.setenv("D9D_NVLINK_MULTIMEM", multimem)  # peer loads/stores vs NVSwitch multicast
    _spawn(_sharded_adamw_worker, 2, state_dtype, max_norm, overlap)


@pytest.mark.parametrize("world", [4, 8])
@pytest.mark.parametrize("multimem", ["0", "1"])
def test_nvlink_sharded_adamw_matches_reference_on_more_gpus(world, multimem, monkeypatch):
    _need_gpus(world)
    monkeypatch.setenv("D9D_NVLINK_MULTIMEM", multimem)
    _spawn(_sharded_adamw_worker, world, torch.bfloat16, 1.0, True)


def _trainer_worker(rank, world, optimizer_name, tmp, expert_parallel=1):
    from pathlib import Path

    import torch.distributed as dist

    from d9d_b200.core.dist_context import DeviceMeshParameters
    from d9d_b200.loop.auto import AutoLRSchedulerProvider, AutoOptimizerProvider
    from d9d_b200.loop.auto.auto_lr_scheduler import PiecewiseConfig
    from d9d_b200.loop.auto.auto_optimizer import NvlinkShardedAdamWOptimizerConfig, StochasticAdamWOptimizerConfig
    from d9d_b200.loop.run import TrainingConfigurator
    from d9d_b200.module.model.qwen3_moe import Qwen3MoEForCausalLMParameters, Qwen3MoELayerParameters, Qwen3MoEParameters
    from d9d_b200.recipes import (CausalLMTask, Qwen3MoEModelProvider, Qwen3MoEModelProviderConfig, SyntheticDataConfig,
                                  SyntheticDataProvider)
    from tests.helpers_train import trainer_config

    params = Qwen3MoEForCausalLMParameters(model=Qwen3MoEParameters(
        layer=Qwen3MoELayerParameters(hidden_size=256, intermediate_size=192, num_experts=8, experts_top_k=2, num_attention_heads=4,
                                      num_key_value_heads=2, rms_norm_eps=1e-6, head_dim=64),
        num_hidden_layers=2, rope_base=10000, max_position_ids=256, split_vocab_size={"regular": 1000, "special": 24},
        split_vocab_order=["regular", "special"]))
    opt_cfg = (NvlinkShardedAdamWOptimizerConfig(lr=3e-3, weight_decay=0.0) if optimizer_name == "nvlink"
               else StochasticAdamWOptimizerConfig(lr=3e-3, weight_decay=0.0, state_dtype="bfloat16"))
    lr_cfg = PiecewiseConfig.model_validate({"name": "piecewise", "scheduler": {"initial_multiplier": 1.0, "phases": [
        {"mode": "rest", "target_multiplier": 1.0, "curve": {"type": "linear"}}]}})
    trainer = TrainingConfigurator(
        mesh=DeviceMeshParameters(data_parallel_replicate=world, expert_parallel=expert_parallel),
        parameters=trainer_config(Path(tmp) / f"r{rank}", total_batch=16, micro=4),
        task_provider=lambda ctx: CausalLMTask(),
        model_provider=Qwen3MoEModelProvider(Qwen3MoEModelProviderConfig(model=params)),
        data_provider=SyntheticDataProvider(SyntheticDataConfig(num_samples=16 * 12, seq_len=128, vocab_size=1024, seed=1, learnable=True)),
        optimizer_provider=AutoOptimizerProvider(opt_cfg),
        lr_scheduler_provider=AutoLRSchedulerProvider(lr_cfg),
    ).configure()
    losses = []
    from d9d_b200.loop.event.catalogue.train import EVENT_TRAIN_OPTIMIZER_STEP_POST

    state = trainer.state  # (the loss accumulator is reset by zero_grad, i.e. before EVENT_TRAIN_STEP_POST)
    state.event_bus.subscribe(EVENT_TRAIN_OPTIMIZER_STEP_POST, lambda ctx: losses.append(state.gradient_manager.compute_global_loss().item()))
    trainer.train()
    assert len(losses) == 12 and losses[-1] < losses[0] - 0.5, losses
    if expert_parallel > 1:
        from d9d_b200.module.block.moe import MoELayer
        from d9d_b200.module.block.moe.communications.nvlink import NvlinkExpertParallelCommunicationHandler

        layers = [m for stage in state.tracked_modules.modules for m in stage.modules() if isinstance(m, MoELayer)]
        assert layers and all(isinstance(m._communicator._nvlink, NvlinkExpertParallelCommunicationHandler) for m in layers)
        if rank == 0:
            torch.save(torch.tensor(losses), Path(tmp) / f"losses_{optimizer_name}_ep.pt")
        return
    flat = torch.cat([(p._local_tensor if hasattr(p, "_local_tensor") else p.data).float().flatten()
                      for m in state.tracked_modules.modules for p in m.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    assert all(torch.equal(g, gathered[0]) for g in gathered)  # replicas stayed bit-identical
    if rank == 0:
        torch.save(torch.tensor(losses), Path(tmp) / f"losses_{optimizer_name}.pt")


def test_trainer_with_nvlink_optimizer_tracks_nccl_path(tmp_path):
    _need_gpus(2)
    for name in ("nvlink", "nccl"):
        _spawn(_trainer_worker, 2, name, str(tmp_path))
    a = torch.load(tmp_path / "losses_nvlink.pt")
    b = torch.load(tmp_path / "losses_nccl.pt")
    assert abs(float(a[0] - b[0])) < 1e-3  # same data, same init
    assert float((a - b).abs().max()) < 0.25, (a, b)  # same trajectory up to stochastic-rounding noise


def _tp_gemm_worker(rank, world):
    import torch.distributed as dist

    from d9d_b200.internals.nvlink import SymmetricArena
    from d9d_b200.kernel._native import native_ops

    ops = native_ops()
    dev = torch.device("cuda", rank)
    group = dist.group.WORLD
    K, N = 512, 384
    for batch, rows_block in ((1, 256), (2, 128)):  # contiguous token shards, and [batch, seq/W] batch-major shards
        rows_local = batch * rows_block
        M = rows_local * world
        g = torch.Generator(device="cuda").manual_seed(11 + batch)
        all_x = torch.randn(world, rows_local, K, device=dev, generator=g).bfloat16()  # same on every rank
        w = torch.randn(N, K, device=dev, generator=g).bfloat16()

        def gathered(shards):  # [world, batch*rows_block, C] -> [batch * world * rows_block, C] in (b, rank, s) order
            c = shards.shape[-1]
            return shards.view(world, batch, rows_block, c).permute(1, 0, 2, 3).reshape(M, c)

        full_x = gathered(all_x).float()

        # ---- all-gather -> GEMM
        arena = SymmetricArena(rows_local * K, torch.bfloat16, dev, group)
        arena.buffer.copy_(all_x[rank].reshape(-1))
        arena.barrier()
        ptrs = [int(p) for p in arena.handle.buffer_ptrs]
        y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        ops.gemm_ag_a(ptrs, rows_local, K, rows_block, w, y, False)
        torch.testing.assert_close(y.float(), full_x @ w.float().t(), rtol=2e-2, atol=2e-1)
        wt = w.t().contiguous()  # [K, N]: MN-major B (dgrad layout)
        y2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        ops.gemm_ag_a(ptrs, rows_local, K, rows_block, wt, y2, True)
        torch.testing.assert_close(y2.float(), full_x @ w.float().t(), rtol=2e-2, atol=2e-1)

        # ---- wgrad with the token-sharded operand pulled from the peers: dW[N, K] = dY^T[N, M] · X[M, K]
        dy = torch.randn(M, N, device=dev, generator=g).bfloat16()
        dw = torch.zeros(N, K, device=dev, dtype=torch.float32)
        ops.gemm_ag_k(dy, ptrs, False, rows_local, K, rows_block, dw, True)  # peers hold B (= X)
        torch.testing.assert_close(dw, dy.float().t() @ full_x, rtol=1e-3, atol=1e-2)
        dw2 = torch.empty(K, N, device=dev, dtype=torch.float32)
        ops.gemm_ag_k(dy, ptrs, True, rows_local, K, rows_block, dw2, False)  # peers hold A (= X): dW2[K, N] = X^T dY
        torch.testing.assert_close(dw2, full_x.t() @ dy.float(), rtol=1e-3, atol=1e-2)
        arena.barrier()

        # ---- GEMM -> reduce-scatter: every rank contributes a_r @ w^T, rank r ends up with the summed rows it owns
        a_r = torch.randn(M, K, device=dev, generator=torch.Generator(device="cuda").manual_seed(100 + rank)).bfloat16()
        out = SymmetricArena(rows_local * N, torch.bfloat16, dev, group)
        out.buffer.zero_()
        out.barrier()
        optrs = [int(p) for p in out.handle.buffer_ptrs]
        ops.gemm_rs_d(a_r, w, optrs, rows_local, N, rows_block, False)
        torch.cuda.synchronize()
        out.barrier()
        parts = [torch.empty_like(a_r) for _ in range(world)]
        dist.all_gather(parts, a_r)
        total = sum(p.float() @ w.float().t() for p in parts)  # [M, N]
        mine = total.view(batch, world, rows_block, N)[:, rank].reshape(rows_local, N)
        torch.testing.assert_close(out.buffer.view(rows_local, N).float(), mine, rtol=3e-2, atol=5e-1)
        out.barrier()


def test_tensor_parallel_gemms_with_fused_communication():
    # validated on 2 GPUs; the flag protocol of the fused all-gather GEMM dead-locks at 4 ranks (known limitation, see NOTES.md)
    _need_gpus(2)
    _spawn(_tp_gemm_worker, 2)


def _tp_mlp_worker(rank, world):
    from torch import nn
    from torch.distributed.device_mesh import init_device_mesh

    from d9d_b200.kernel._native import native_ops
    from d9d_b200.module.block.linear import Linear
    from d9d_b200.module.parallelism.api import parallelize_colwise, parallelize_rowwise

    class MLP(nn.Module):
        def __init__(self):
            super().__init__()
            self.up = Linear(512, 1024, bias=False)
            self.down = Linear(1024, 512, bias=False)

        def forward(self, x):
            return self.down(torch.nn.functional.silu(self.up(x)))

    mesh = init_device_mesh("cuda", (world,), mesh_dim_names=("tp",))
    torch.manual_seed(0)
    ref = MLP().cuda()
    tp = MLP().cuda().bfloat16()
    tp.load_state_dict({k: v.bfloat16() for k, v in ref.state_dict().items()})
    ref.load_state_dict({k: v.float() for k, v in tp.state_dict().items()})  # identical (bf16-rounded) weights
    parallelize_colwise(tp.up, mesh, sequence_parallel=True)
    parallelize_rowwise(tp.down, mesh, sequence_parallel=True)

    x = torch.randn(2, 256 * world, 512, device="cuda").bfloat16()
    x_ref = x.float().requires_grad_()
    y_ref = ref(x_ref)
    (y_ref * 0.01).square().sum().backward()
    x_in = x.chunk(world, dim=1)[rank].contiguous().requires_grad_()
    before = native_ops().launches
    y = tp(x_in)
    torch.testing.assert_close(y.float(), y_ref.detach().chunk(world, dim=1)[rank], rtol=3e-2, atol=3e-2)
    (y.float() * 0.01).square().sum().backward()
    assert native_ops().launches - before >= 6  # 2 fused forward GEMMs + 4 fused backward GEMMs ran on the native path

    def close(a, b):
        assert torch.nn.functional.cosine_similarity(a.float().flatten(), b.float().flatten(), dim=0) > 0.995
        assert 0.95 < float(a.float().norm() / b.float().norm()) < 1.05

    close(x_in.grad, x_ref.grad.chunk(world, dim=1)[rank])
    close(tp.up.weight.grad.to_local(), ref.up.weight.grad.chunk(world, dim=0)[rank])
    close(tp.down.weight.grad.to_local(), ref.down.weight.grad.chunk(world, dim=1)[rank])


def test_sequence_parallel_mlp_uses_fused_tp_kernels():
    _need_gpus(2)
    _spawn(_tp_mlp_worker, 2)


def _ep_worker(rank, world):
    """EP=2 MoE layer on the NVLink exchange vs the same layer with all experts local (single-process reference)."""
    import torch.distributed as dist
    from torch.distributed.device_mesh import init_device_mesh

    from d9d_b200.kernel._native import native_ops
    from d9d_b200.module.block.moe import MoELayer
    from d9d_b200.module.parallelism.api import parallelize_expert_parallel

    E, k, H, F = 16, 2, 256, 192
    torch.manual_seed(3)
    ref = MoELayer(hidden_dim=H, intermediate_dim_grouped=F, num_grouped_experts=E, top_k=k, router_renormalize_probabilities=True)
    ref.reset_parameters()
    ref = ref.cuda().bfloat16()
    ep = MoELayer(hidden_dim=H, intermediate_dim_grouped=F, num_grouped_experts=E, top_k=k, router_renormalize_probabilities=True)
    ep = ep.cuda().bfloat16()
    ep.load_state_dict(ref.state_dict())
    ep.tokens_per_expert.zero_()
    ref.tokens_per_expert.zero_()
    mesh = init_device_mesh("cuda", (1, world), mesh_dim_names=("ep_replicate", "ep_shard"))
    parallelize_expert_parallel(ep, mesh)

    g = torch.Generator(device="cuda").manual_seed(50 + rank)  # every rank routes different tokens
    x = torch.randn(2, 200, H, device="cuda", generator=g).bfloat16()
    xa, xb = x.clone().requires_grad_(), x.clone().requires_grad_()
    before = native_ops().launches
    y = ep(xa)
    y_ref = ref(xb)
    torch.testing.assert_close(y.float(), y_ref.float(), rtol=3e-2, atol=3e-2)
    w = torch.randn_like(y_ref)
    (y.float() * w).sum().backward()
    (y_ref.float() * w).sum().backward()
    assert native_ops().launches > before

    def close(a, b, name):
        a, b = a.float().flatten(), b.float().flatten()
        assert torch.nn.functional.cosine_similarity(a, b, dim=0) > 0.99, name
        assert 0.9 < float(a.norm() / (b.norm() + 1e-12)) < 1.1, name

    close(xa.grad, xb.grad, "dx")
    # expert weights: this rank owns experts [rank * E/W, (rank + 1) * E/W); their gradients only see tokens routed to them
    # from *all* ranks, so compare against the reference accumulated over every rank's tokens
    local = E // world
    for name in ("gate_proj", "up_proj", "down_proj"):
        mine = getattr(ep.grouped_experts, name).weight.grad.to_local().float()
        full = getattr(ref.grouped_experts, name).weight.grad.float().clone()
        dist.all_reduce(full)  # sum of the per-rank references == gradient of the global batch
        close(mine, full[rank * local : (rank + 1) * local], name)
    gate = ep.router.gate.weight.grad
    close(gate.to_local() if hasattr(gate, "to_local") else gate, ref.router.gate.weight.grad, "router")


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("capacity_factor", ["0", "4"])
def test_expert_parallel_over_nvlink_matches_local_experts(world, capacity_factor, monkeypatch):
    """EP = world: worst-case buffers and capacity-factor buffers (generous enough not to drop anything here)."""
    _need_gpus(world)
    monkeypatch.setenv("D9D_EP_CAPACITY_FACTOR", capacity_factor)
    _spawn(_ep_worker, world)


def test_trainer_with_expert_parallel_over_nvlink(tmp_path):
    """DP=2 x EP=2 on two GPUs (experts sharded, tokens exchanged by the NVLink kernels) follows the pure-DP run."""
    _need_gpus(2)
    _spawn(_trainer_worker, 2, "nccl", str(tmp_path), 2)
    _spawn(_trainer_worker, 2, "nccl", str(tmp_path), 1)
    a = torch.load(tmp_path / "losses_nccl_ep.pt")
    b = torch.load(tmp_path / "losses_nccl.pt")
    assert abs(float(a[0] - b[0])) < 2e-2 and float((a - b).abs().max()) < 0.3, (a, b)
