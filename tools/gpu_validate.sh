#!/usr/bin/env bash
# One-shot GPU validation for a box with very little time: every step runs under its own timeout in its own process (a
# trapping kernel cannot poison the next step) and logs to gpurun_out/val/.  Usage: tools/gpu_validate.sh tests|bench|all
set -u
phase="${1:-all}"
out=gpurun_out/val
mkdir -p "$out"
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 ))s] $*" | tee -a "$out/timeline.log"; }

run_bg() {  # name timeout cmd...
  local name="$1" to="$2"; shift 2
  ( timeout "$to" "$@" > "$out/$name.log" 2>&1; echo "rc=$?" >> "$out/$name.log" ) &
}

if [[ "$phase" == "tests" || "$phase" == "all" ]]; then
  python -c "import torch; print(torch.cuda.get_device_name(0))" > "$out/device.log" 2>&1
  stamp "torch imported"
  run_bg t_ops 150 python -m pytest tests/test_ops_gpu.py -q -p no:cacheprovider
  run_bg t_pair 150 python -m pytest tests/test_zz_pair_gpu.py -q -p no:cacheprovider
  run_bg t_fp8 150 python -m pytest tests/test_zz_fp8_gpu.py -q -p no:cacheprovider
  run_bg t_model 200 python -m pytest tests/test_model_gpu.py tests/test_moe_experts_gpu.py tests/test_flash_attn_gpu.py -q -p no:cacheprovider
  wait
  stamp "tests done"
  for v in single pair fp8 mxfp8; do
    timeout 90 python benchmarks/bench_gemm_variants.py --variants cublas,$v --iters 5 --out "$out/variants_$v.json" > "$out/variants_$v.log" 2>&1
    stamp "variants $v rc=$?"
  done
fi

if [[ "$phase" == "bench" || "$phase" == "all" ]]; then
  timeout 200 python bench.py --steps 4 --warmup 3 > "$out/bench_default.log" 2>&1
  stamp "bench default rc=$?"
  D9D_GEMM_PAIR=1 timeout 160 python bench.py --steps 4 --warmup 3 --no-e2e > "$out/bench_pair.log" 2>&1
  stamp "bench pair rc=$?"
fi
tail -n 3 "$out"/t_*.log 2>/dev/null | tail -n 40
grep -h '"metric"' "$out"/bench_*.log 2>/dev/null | cut -c1-400
stamp "done"
