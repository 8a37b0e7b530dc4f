"""Native tcgen05 flash-attention forward vs PyTorch SDPA (cuDNN) forward, flagship shapes."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from d9d_b200.kernel._native import native_ops

ops = native_ops()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timeit(fn, iters=20):
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]

res = []
for (B, S, Hq, Hk, D, causal) in [(8, 2048, 16, 4, 128, True), (8, 2048, 16, 4, 128, False), (2, 8192, 32, 8, 128, True), (8, 2048, 16, 4, 64, True)]:
    q = torch.randn(B, S, Hq, D, device="cuda").bfloat16(); k = torch.randn(B, S, Hk, D, device="cuda").bfloat16(); v = torch.randn_like(k)
    flops = 4.0 * B * Hq * S * S * D * (0.5 if causal else 1.0)
    t_native = timeit(lambda: ops.flash_attn_fwd(q, k, v, D ** -0.5, causal))
    qh, kh, vh = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
    with torch.no_grad():
        t_sdpa = timeit(lambda: F.scaled_dot_product_attention(qh, kh, vh, is_causal=causal, enable_gqa=True))
    res.append({"B": B, "S": S, "Hq": Hq, "Hk": Hk, "D": D, "causal": causal, "native_ms": t_native, "sdpa_ms": t_sdpa,
                "native_tflops": flops / t_native / 1e9, "sdpa_tflops": flops / t_sdpa / 1e9})
    print(res[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/bench_attention.json", "w"), indent=1)
