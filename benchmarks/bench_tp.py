"""Fused all-gather->GEMM / GEMM->reduce-scatter kernels vs NCCL collective + GEMM (run under torchrun).

Shapes: one Llama-3-8B MLP (hidden 4096, ffn 14336) over `tokens` tokens, tensor parallel over all ranks, sequence
parallel activations.  Times are CUDA-event medians, max over ranks.
"""

from __future__ import annotations

import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=8192)
    ap.add_argument("--hidden", type=int, default=4096)
    ap.add_argument("--ffn", type=int, default=14336)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default="gpurun_out/bench_tp.json")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    from d9d_b200.kernel._native import native_ops
    from d9d_b200.kernel.tp.fused import TensorParallelWorkspace

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    ops = native_ops()
    ws = TensorParallelWorkspace.for_group(dist.group.WORLD)
    T, H, F = args.tokens, args.hidden, args.ffn
    Tl, Fl = T // world, F // world
    bf = dict(device=dev, dtype=torch.bfloat16)
    x_shard = torch.randn(Tl, H, **bf)
    w_up = torch.randn(Fl, H, **bf) * 0.02     # column parallel
    w_down = torch.randn(H, Fl, **bf) * 0.02   # row parallel
    h = torch.randn(T, Fl, **bf)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn):
        ts = []
        for _ in range(args.iters + 3):
            flush.zero_()
            dist.barrier()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        t = torch.tensor([sorted(ts[3:])[len(ts[3:]) // 2]], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    y = torch.empty(T, Fl, **bf)
    gathered = torch.empty(T, H, **bf)

    ones = torch.ones(world, dtype=torch.int32, device=dev)

    def ag_gemm_fused():
        _, ptrs = ws.stage("ag_in", x_shard)
        ops.gemm_ag_a(ptrs, Tl, H, Tl, w_up, y, False)

    def ag_gemm_pipelined(spare=0):
        ws.gather_async("ag_in", x_shard, Tl, consumer=lambda g2, flags: ops.gemm_wait_a(g2, flags, rank, Tl, w_up, y, False, spare))
        ws.join()

    def concurrent_independent():  # gather chain and an ungated GEMM at the same time (no data dependency)
        ws.gather_async("ag_in", x_shard, Tl, consumer=lambda g2, flags: ops.gemm_wait_a(gathered, ones, rank, Tl, w_up, y, False, 0))
        ws.join()

    def wait_gemm_only():  # flag-gated kernel with every flag already raised
        ops.gemm_wait_a(gathered, ones, rank, Tl, w_up, y, False)

    def _unused():
        pass

    def gather_only():
        ws.gather_async("ag_in", x_shard, Tl)
        ws.join()

    def ag_gemm_nccl():
        dist.all_gather_into_tensor(gathered, x_shard)
        ops.gemm(gathered, w_up, y, False, False, False)

    out_shard = torch.empty(Tl, H, **bf)
    partial = torch.empty(T, H, **bf)

    def gemm_rs_fused():
        arena, ptrs = ws.zeroed("rs_out", Tl * H, dev)
        ops.gemm_rs_d(h, w_down, ptrs, Tl, H, Tl, False)
        arena.barrier()
        out_shard.copy_(arena.buffer[: Tl * H].view(Tl, H))

    def gemm_rs_nccl():
        ops.gemm(h, w_down, partial, False, False, False)
        dist.reduce_scatter_tensor(out_shard, partial)

    def gemm_only_up():
        ops.gemm(gathered, w_up, y, False, False, False)

    def gemm_only_down():
        ops.gemm(h, w_down, partial, False, False, False)

    res = {"world": world, "tokens": T, "hidden": H, "ffn": F,
           "ag_gemm_fused_ms": timed(ag_gemm_fused), "ag_gemm_pipelined_ms": timed(ag_gemm_pipelined), "ag_gemm_pipelined_spare8_ms": timed(lambda: ag_gemm_pipelined(8)),
           "max_connections_env": os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"),
           "wait_gemm_only_ms": timed(wait_gemm_only), "gather_only_ms": timed(gather_only), "concurrent_independent_ms": timed(concurrent_independent), "ag_gemm_nccl_ms": timed(ag_gemm_nccl), "gemm_up_only_ms": timed(gemm_only_up),
           "gemm_rs_fused_ms": timed(gemm_rs_fused), "gemm_rs_nccl_ms": timed(gemm_rs_nccl), "gemm_down_only_ms": timed(gemm_only_down)}
    flops = 2.0 * T * H * Fl
    for k in ("ag_gemm_pipelined_ms", "ag_gemm_fused_ms", "ag_gemm_nccl_ms", "gemm_up_only_ms", "gemm_rs_fused_ms", "gemm_rs_nccl_ms", "gemm_down_only_ms"):
        res[k.replace("_ms", "_tflops")] = flops / res[k] / 1e9
    if rank == 0:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        json.dump(res, open(args.out, "w"), indent=1)
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
