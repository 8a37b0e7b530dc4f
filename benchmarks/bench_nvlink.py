"""Micro-benchmark of the NVLink data-parallel optimizer kernels (run under torchrun, one rank per GPU).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 benchmarks/bench_nvlink.py --numel 2730000000
"""

from __future__ import annotations

import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--numel", type=int, default=2_730_000_000)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--out", default="gpurun_out/bench_nvlink.json")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    from d9d_b200.kernel._native import native_ops
    from d9d_b200.optim.nvlink import NvlinkShardedAdamW

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    ops = native_ops()

    p = torch.nn.Parameter(torch.randn(args.numel // 4096, 4096, device=dev).bfloat16())
    opt = NvlinkShardedAdamW([p], dist.group.WORLD, lr=1e-3, max_norm=1.0)
    n = opt._numel
    begin, end = opt._owned_range(0)
    end = begin + opt._chunk  # one owned chunk
    p.grad.normal_()

    def timed(fn, sync_before=True):
        ts = []
        for _ in range(args.iters):
            if sync_before:
                opt.grad_arena.barrier()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        t = torch.tensor([sorted(ts)[len(ts) // 2]], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    res = {"world": world, "numel": n, "multicast": opt.uses_multicast}
    res["barrier_ms"] = timed(lambda: opt.grad_arena.barrier(), sync_before=False)
    res["reduce_shard_ms"] = timed(lambda: ops.nvl_reduce_shard_(opt.grad_arena.buffer, opt.grad_arena.peer_ptrs_dev,
                                                                 opt.grad_arena.multicast_ptr, begin, end, world, rank, opt._sumsq))
    res["reduce_shard_p2p_ms"] = timed(lambda: ops.nvl_reduce_shard_(opt.grad_arena.buffer, opt.grad_arena.peer_ptrs_dev,
                                                                     0, begin, end, world, rank, opt._sumsq))
    res["adamw_shard_ms"] = timed(lambda: ops.nvl_adamw_shard_(
        opt.param_arena.buffer, opt.grad_arena.buffer, opt.exp_avg, opt.exp_avg_sq, opt.param_arena.peer_ptrs_dev,
        opt.param_arena.multicast_ptr, begin, end, world, rank, 1e-3, 0.9, 0.999, 1e-8, 0.01, 0.1, 0.001, 1, None))
    res["adamw_shard_p2p_ms"] = timed(lambda: ops.nvl_adamw_shard_(
        opt.param_arena.buffer, opt.grad_arena.buffer, opt.exp_avg, opt.exp_avg_sq, opt.param_arena.peer_ptrs_dev,
        0, begin, end, world, rank, 1e-3, 0.9, 0.999, 1e-8, 0.01, 0.1, 0.001, 1, None))
    res["zero_grad_ms"] = timed(lambda: opt.grad_arena.buffer.zero_())
    res["full_step_ms"] = timed(lambda: opt.step(), sync_before=False)
    flat = opt.grad_arena.buffer
    res["nccl_allreduce_fp32_ms"] = timed(lambda: dist.all_reduce(flat))
    res["chunk_numel"] = opt._chunk
    res["chunks_per_rank"] = opt._rows
    shard_bytes = (end - begin) * 4
    res["reduce_shard_GBps_in"] = shard_bytes / res["reduce_shard_ms"] / 1e6
    res["allreduce_busbw_GBps"] = 2 * (world - 1) / world * n * 4 / res["nccl_allreduce_fp32_ms"] / 1e6
    if rank == 0:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
