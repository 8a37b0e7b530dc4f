import os, time, torch
print({k: v for k, v in os.environ.items() if k.startswith(("CUDA", "NCCL", "TORCH", "NVIDIA"))}, flush=True)
dev = torch.device("cuda", 0)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
x = torch.zeros(1, device=dev)
torch.cuda.synchronize()
def run(parallel):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    t0 = time.time()
    if parallel:
        s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s1): torch.cuda._sleep(400_000_000)
        with torch.cuda.stream(s2): torch.cuda._sleep(400_000_000)
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    else:
        torch.cuda._sleep(400_000_000); torch.cuda._sleep(400_000_000)
    host = time.time() - t0
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b), host
print("serial   ms, host-enqueue s:", run(False))
print("parallel ms, host-enqueue s:", run(True))
