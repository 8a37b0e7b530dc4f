"""Kernel-level timeline of one optimizer step of the flagship model (torch.profiler, CUPTI).

    python benchmarks/profile_step.py --layers 16 --out gpurun_out/step_profile.json

Writes the per-kernel totals (sorted by device time), the GPU-busy fraction of the step and the CPU-side launch
count.  Numbers taken under the profiler are for *attribution only* – never bench values.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=16)
    ap.add_argument("--seq-len", type=int, default=2048)
    ap.add_argument("--microbatch", type=int, default=8)
    ap.add_argument("--accum", type=int, default=2)
    ap.add_argument("--out", default="gpurun_out/step_profile.json")
    ap.add_argument("--trace", default=None, help="optional chrome trace path")
    args = ap.parse_args()
    args.warmup, args.steps = 2, 1
    for name, value in (('model', 'example'), ('layout', 'dp'), ('impl', 'own'), ('checkpointing', False), ('dp_impl', 'nvlink'), ('gpus', 1)):
        if not hasattr(args, name):
            setattr(args, name, value)

    import torch
    from torch.profiler import ProfilerActivity, profile

    import bench
    from d9d_b200 import ops
    from d9d_b200.bench_support import TrainStepRunner

    device = torch.device("cuda", 0)
    torch.cuda.set_device(device)
    ops.load()
    runner = TrainStepRunner(args, device, 1, bench.build_model)
    vocab = sum(bench.FLAGSHIP["split_vocab_size"].values())
    gen = torch.Generator(device=device).manual_seed(1)
    batches = [runner.synthetic_batch(vocab, gen) for _ in range(4 * args.accum)]
    for i in range(2):
        runner.step(batches[i * args.accum : (i + 1) * args.accum])
    torch.cuda.synchronize()

    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    runner.step(batches[2 * args.accum : 3 * args.accum])
    e.record()
    torch.cuda.synchronize()
    unprofiled_ms = s.elapsed_time(e)

    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        runner.step(batches[3 * args.accum : 4 * args.accum])
        torch.cuda.synchronize()
    if args.trace:
        prof.export_chrome_trace(args.trace)

    kernels = defaultdict(lambda: [0, 0.0])
    intervals = []
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            k = kernels[ev.name]
            k[0] += 1
            k[1] += ev.device_time
            intervals.append((ev.time_range.start, ev.time_range.end))
    intervals.sort()
    busy, cur_s, cur_e = 0.0, None, None
    for a, b in intervals:
        if cur_e is None or a > cur_e:
            if cur_e is not None:
                busy += cur_e - cur_s
            cur_s, cur_e = a, b
        else:
            cur_e = max(cur_e, b)
    if cur_e is not None:
        busy += cur_e - cur_s
    span = (intervals[-1][1] - intervals[0][0]) if intervals else 0.0
    total = sum(v[1] for v in kernels.values())
    rows = sorted(({"name": n[:140], "calls": c, "total_us": round(t, 1), "pct": round(100 * t / max(total, 1e-9), 2)}
                   for n, (c, t) in kernels.items()), key=lambda r: -r["total_us"])
    out = {
        "config": {"layers": args.layers, "seq_len": args.seq_len, "microbatch": args.microbatch, "accum": args.accum},
        "unprofiled_step_ms": unprofiled_ms,
        "profiled_span_ms": span / 1e3,
        "gpu_busy_ms": busy / 1e3,
        "gpu_busy_fraction_of_span": busy / max(span, 1e-9),
        "sum_kernel_ms": total / 1e3,
        "num_device_events": len(intervals),
        "kernels": rows[:80],
    }
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "kernels"}))
    for r in rows[:45]:
        print(f"{r['pct']:6.2f}% {r['total_us']/1e3:9.3f} ms {r['calls']:6d}  {r['name'][:110]}")


if __name__ == "__main__":
    main()
