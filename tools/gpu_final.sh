#!/usr/bin/env bash
# Final validation on a 2-GPU box: the whole single-GPU suite + smoke on GPU 0, the data-parallel bench on 2 GPUs (NVLink
# optimizer + CTA-pair GEMMs + chunked CE), a quick 1-GPU bench and the fp8 experiment.  Ordered by priority; every step has
# its own timeout.
set -u
out=gpurun_out/final
mkdir -p "$out"
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 ))s] $*" | tee -a "$out/timeline.log"; }
( CUDA_VISIBLE_DEVICES=0 timeout 170 python -m pytest tests -m gpu -q -p no:cacheprovider -x > "$out/pytest_gpu.log" 2>&1; echo "rc=$?" >> "$out/pytest_gpu.log" ) &
( CUDA_VISIBLE_DEVICES=1 timeout 120 python __graft_entry__.py smoke > "$out/smoke.log" 2>&1; echo "rc=$?" >> "$out/smoke.log" ) &
wait
stamp "pytest + smoke done"; tail -n 2 "$out/pytest_gpu.log"; tail -n 2 "$out/smoke.log"
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 4 --warmup 3 --no-e2e > "$out/bench_2gpu.log" 2>&1
stamp "bench 2gpu rc=$?"
CUDA_VISIBLE_DEVICES=0 timeout 100 python bench.py --steps 6 --warmup 3 --no-e2e > "$out/bench_1gpu.log" 2>&1
stamp "bench 1gpu rc=$?"
CUDA_VISIBLE_DEVICES=0 timeout 100 python bench.py --steps 4 --warmup 3 --no-e2e --fp8-dense > "$out/bench_1gpu_fp8.log" 2>&1
stamp "bench 1gpu fp8 rc=$?"
grep -h -o '"ms_per_step": [0-9.]*\|"final_loss": [0-9.]*\|"n_gpus": [0-9]*' "$out"/bench_*.log | tr '\n' ' '; echo
stamp done
