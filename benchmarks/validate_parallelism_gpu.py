"""GPU validation of the parallelism features that so far have only run on gloo (context parallel attention, whole-model
tensor / context parallel plans).  Not part of the pytest suite: run it on a multi-GPU box, e.g.

    torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node 2 benchmarks/validate_parallelism_gpu.py
    torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node 4 benchmarks/validate_parallelism_gpu.py --meshes cps4 dpr2_tp2 dps2_tp2 cpr2_tp2

Every check compares against the same computation on one GPU (bf16 on the native kernels, so tolerances are loose:
gradients are compared by direction and norm).  Prints one line per check and exits non-zero on the first failure.
"""

from __future__ import annotations

import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

DEVICE, DTYPE = "cuda", torch.bfloat16  # ``--cpu`` (logic dry run on gloo) switches to cpu / fp32

MESHES = {
    # 2 GPUs
    "dpr2": dict(data_parallel_replicate=2), "dps2": dict(data_parallel_shard=2), "dpr2_ep2": dict(data_parallel_replicate=2, expert_parallel=2),
    "cps2": dict(context_parallel_shard=2), "tp2": dict(tensor_parallel=2),
    # 4 GPUs
    "dpr4": dict(data_parallel_replicate=4), "dps4": dict(data_parallel_shard=4), "dpr2_dps2": dict(data_parallel_replicate=2, data_parallel_shard=2),
    "dpr4_ep4": dict(data_parallel_replicate=4, expert_parallel=4), "dps4_ep2": dict(data_parallel_shard=4, expert_parallel=2),
    "cps4": dict(context_parallel_shard=4), "dpr2_tp2": dict(data_parallel_replicate=2, tensor_parallel=2),
    "dps2_tp2": dict(data_parallel_shard=2, tensor_parallel=2), "cpr2_tp2": dict(context_parallel_replicate=2, tensor_parallel=2),
    "dpr2_cps2": dict(data_parallel_replicate=2, context_parallel_shard=2),
    # 8 GPUs: the single-stage meshes of the reference's distributed tier (test/d9d_test/modules/model/meshes.py) + TP / CP
    "dpr8": dict(data_parallel_replicate=8), "dps8": dict(data_parallel_shard=8),
    "dpr2_dps4": dict(data_parallel_replicate=2, data_parallel_shard=4),
    "dpr8_ep2": dict(data_parallel_replicate=8, expert_parallel=2), "dps8_ep2": dict(data_parallel_shard=8, expert_parallel=2),
    "dpr2_dps4_ep2": dict(data_parallel_replicate=2, data_parallel_shard=4, expert_parallel=2),
    "dpr8_ep8": dict(data_parallel_replicate=8, expert_parallel=8),
    "dps4_tp2": dict(data_parallel_shard=4, tensor_parallel=2),
}


def _close(ours: torch.Tensor, ref: torch.Tensor, what: str, cos_tol: float = 2e-3, norm_tol: float = 3e-2) -> None:
    a, b = ours.flatten().double(), ref.flatten().double()
    cos = float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30))
    ratio = float(a.norm() / (b.norm() + 1e-30))
    if not (cos > 1 - cos_tol and abs(ratio - 1) < norm_tol):
        raise AssertionError(f"{what}: cos={cos:.6f} norm ratio={ratio:.4f}")


def check_attention(mode: str, layout_name: str, causal: bool) -> None:
    from d9d_b200.kernel.context_parallel import ContextParallelLayout, local_sequence_indices, ring_attention, shard_sequence, ulysses_attention
    from d9d_b200.kernel.flash_attn import flash_attn_func

    group, world, rank = dist.group.WORLD, dist.get_world_size(), dist.get_rank()
    layout = ContextParallelLayout(layout_name)
    torch.manual_seed(0)
    batch, seq, heads, kv_heads, dim = (2, 256 * world, 8, 4, 128) if DEVICE == "cuda" else (1, 16 * world, 8, 4, 8)
    q, k, v, w = (torch.randn(batch, seq, h, dim, device=DEVICE, dtype=DTYPE) for h in (heads, kv_heads, kv_heads, heads))
    ref_in = [t.clone().requires_grad_() for t in (q, k, v)]
    ref_out, _ = flash_attn_func(*ref_in, causal=causal)
    (ref_out.float() * w.float()).sum().backward()
    local = [shard_sequence(t, 1, world, rank, layout).clone().requires_grad_() for t in (q, k, v)]
    positions = torch.stack([local_sequence_indices(seq, world, r, layout) for r in range(world)])
    if mode == "ring":
        out = ring_attention(*local, group, positions, causal=causal, mask_cache={})
    else:
        out = ulysses_attention(*local, group, lambda a, b, c: flash_attn_func(a, b, c, causal=causal)[0], positions=positions.reshape(-1).to(DEVICE))
    _close(out.float(), shard_sequence(ref_out.detach(), 1, world, rank, layout).float(), "output")
    (out.float() * shard_sequence(w, 1, world, rank, layout).float()).sum().backward()
    for name, mine, full in zip("qkv", local, ref_in):
        _close(mine.grad.float(), shard_sequence(full.grad, 1, world, rank, layout).float(), f"d{name}")


def check_model(mesh_name: str, moe: bool) -> None:
    from torch.distributed.tensor import DTensor

    from d9d_b200.core.dist_context import BATCH_DOMAIN, DeviceMeshParameters
    from d9d_b200.dataset import shard_batch_along_sequence
    from d9d_b200.internals.grad_sync import GradientSynchronizer
    from d9d_b200.module.block.hidden_states_aggregator import HiddenStatesAggregationMode
    from d9d_b200.pipelining.api import PipelineStageInfo

    small = DEVICE != "cuda"
    hidden, vocab, seq = (64, 128, 32) if small else (512, 4096, 1024)

    def build():
        torch.manual_seed(5)
        if moe:
            from d9d_b200.module.model.qwen3_moe import Qwen3MoEForCausalLM as Cls, Qwen3MoEForCausalLMParameters as P, Qwen3MoELayerParameters as L, Qwen3MoEParameters as B
            layer = L(hidden_size=hidden, intermediate_size=hidden // 2, num_experts=16, experts_top_k=2, num_attention_heads=8, num_key_value_heads=4, rms_norm_eps=1e-6, head_dim=hidden // 8)
        else:
            from d9d_b200.module.model.qwen3_dense import Qwen3DenseForCausalLM as Cls, Qwen3DenseForCausalLMParameters as P, Qwen3DenseLayerParameters as L, Qwen3DenseParameters as B
            layer = L(hidden_size=hidden, intermediate_size=2 * hidden, num_attention_heads=8, num_key_value_heads=4, rms_norm_eps=1e-6, head_dim=hidden // 8)
        params = P(model=B(layer=layer, num_hidden_layers=2, rope_base=10000, max_position_ids=2048, split_vocab_size={"text": vocab}, split_vocab_order=["text"]))
        with torch.device(DEVICE):
            model = Cls(params, PipelineStageInfo(0, 1), HiddenStatesAggregationMode.no, False).to(DTYPE)
        model.reset_parameters()
        return model

    def batch(i):
        g = torch.Generator().manual_seed(900 + i)
        ids = torch.randint(0, vocab, (2, seq), generator=g).to(DEVICE)
        labels = torch.randint(0, vocab, (2, seq), generator=g).to(DEVICE)
        return ids, labels, torch.arange(seq, device=DEVICE)[None].expand(2, -1).contiguous()

    ctx = DeviceMeshParameters(**MESHES[mesh_name]).build()
    model = build()
    if moe:
        from d9d_b200.module.parallelism.model.qwen3_moe import parallelize_qwen3_moe_for_causal_lm as plan
    else:
        from d9d_b200.module.parallelism.model.qwen3_dense import parallelize_qwen3_dense_for_causal_lm as plan
    plan(ctx, model, PipelineStageInfo(0, 1))
    params = list(model.parameters())
    for p in params:
        p.grad_dtype = torch.float32
    sync = GradientSynchronizer([params], bucket_size_mb=64, require_accumulations=1)
    sync.bind()
    mesh = ctx.mesh_for(BATCH_DOMAIN)
    ids, labels, pos = shard_batch_along_sequence(batch(mesh["dp"].get_local_rank()), ctx)
    model(input_ids=ids.contiguous(), position_ids=pos.contiguous(), labels=labels.contiguous())["logps"].sum().backward()
    sync.wait()
    ref = build()
    for p in ref.parameters():
        p.grad_dtype = torch.float32
    for b in range(mesh["dp"].size()):
        ids, labels, pos = batch(b)
        ref(input_ids=ids, position_ids=pos, labels=labels)["logps"].sum().backward()
    for (name, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        got = p.grad.full_tensor() if isinstance(p.grad, DTensor) else p.grad
        _close(got.float(), q.grad.float(), f"[{mesh_name}] {name}", cos_tol=2e-2, norm_tol=5e-2)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--meshes", nargs="*", default=None)
    ap.add_argument("--skip-attention", action="store_true")
    ap.add_argument("--cpu", action="store_true", help="dry run of the script's logic on gloo / cpu / fp32 with tiny shapes")
    args = ap.parse_args()
    global DEVICE, DTYPE
    if args.cpu:
        DEVICE, DTYPE = "cpu", torch.float32
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    world, rank = dist.get_world_size(), dist.get_rank()
    failures = 0

    def run(label, fn, *a):
        nonlocal failures
        try:
            fn(*a)
            status = "ok"
        except Exception as exc:  # noqa: BLE001
            status, failures = f"FAILED: {exc!r}", failures + 1
        flags = [None] * world
        dist.all_gather_object(flags, status)
        if rank == 0:
            bad = [f"rank {i}: {s}" for i, s in enumerate(flags) if s != "ok"]
            print(f"{label:60s} {'ok' if not bad else bad}", flush=True)

    for mode in (() if args.skip_attention else ("ulysses", "ring")):
        for layout in ("zigzag", "contiguous"):
            for causal in (True, False):
                run(f"attention {mode} {layout} causal={causal} (world {world})", check_attention, mode, layout, causal)
    meshes = args.meshes if args.meshes is not None else [m for m, kw in MESHES.items() if _size(kw) == world]
    for mesh_name in meshes:
        for moe in (False, True):
            if not moe and "ep" in mesh_name:
                continue  # expert parallelism needs experts
            run(f"model {'moe' if moe else 'dense'} {mesh_name}", check_model, mesh_name, moe)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if failures else 0)


def _size(kwargs: dict) -> int:
    n = 1
    for key, value in kwargs.items():
        if key != "expert_parallel":
            n *= value
    return n


if __name__ == "__main__":
    main()
