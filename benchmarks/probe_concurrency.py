"""Can a small kernel on a side stream run while the persistent flag-gated GEMM spins? (1 GPU probe)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from d9d_b200.kernel._native import native_ops

ops = native_ops()
dev = torch.device("cuda", 0)
T, H, N = 8192, 4096, 7168
a = torch.randn(T, H, device=dev, dtype=torch.bfloat16)
w = torch.randn(N, H, device=dev, dtype=torch.bfloat16)
y = torch.empty(T, N, device=dev, dtype=torch.bfloat16)
side = torch.cuda.Stream()
# load the side-chain kernels up front: with lazy module loading a first-time launch needs the device to drain, which
# dead-locks against a kernel that is spinning on the very flag that launch would raise
_warm = torch.zeros(2, dtype=torch.int32, device=dev); _warm[1:2].fill_(1); torch.cuda._sleep(1000); torch.cuda.synchronize()
print("CUDA_DEVICE_MAX_CONNECTIONS", os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"), flush=True)
for spare in (8, 0):
    for order in ("gemm_first", "side_first"):
        flags = torch.zeros(2, dtype=torch.int32, device=dev)
        flags[0] = 1
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        ready = torch.cuda.Event(); ready.record()
        side_done = torch.cuda.Event()
        def side_work():
            side.wait_event(ready)
            with torch.cuda.stream(side):
                torch.cuda._sleep(200_000)  # ~0.1 ms of spinning kernel
                flags[1:2].fill_(1)
                side_done.record()
        if order == "side_first":
            side_work()
        ops.gemm_wait_a(a, flags, 0, T // 2, w, y, False, spare)  # rank 0 owns rows [0, T/2); the rest waits for flags[1]
        if order == "gemm_first":
            side_work()
        torch.cuda.current_stream().wait_stream(side)
        e.record()
        t0 = time.time()
        while not e.query():
            if side_done.query() and not globals().get("_said"):
                print(f"   side chain finished after {time.time()-t0:.3f}s while the GEMM was still running", flush=True); _said = True
            if time.time() - t0 > 8:
                print(f"spare={spare} {order}: STUCK (flag never observed)", flush=True)
                os._exit(3)
            time.sleep(0.01)
        print(f"spare={spare} {order}: {s.elapsed_time(e):.3f} ms", flush=True)
