#!/usr/bin/env bash
# call 2: verify the test fix, bench the flagship step with / without CTA pairs and with the previous CE scheme, step profile, ncu
set -u
out=gpurun_out/val2
mkdir -p "$out"
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 ))s] $*" | tee -a "$out/timeline.log"; }
timeout 60 python -m pytest tests/test_ops_gpu.py -q -p no:cacheprovider -k "split_blocks or slice_options or accumulate" > "$out/t_ce.log" 2>&1; stamp "ce tests rc=$?"
B="--steps 6 --warmup 3 --no-e2e"
timeout 90 python bench.py $B > "$out/bench_single.log" 2>&1; stamp "bench single rc=$?"
D9D_GEMM_PAIR=1 timeout 90 python bench.py $B > "$out/bench_pair.log" 2>&1; stamp "bench pair rc=$?"
D9D_GEMM_PAIR=1 D9D_CCE_CAT_SPLITS=1 D9D_CCE_CHUNK_BYTES=6442450944 timeout 90 python bench.py $B > "$out/bench_pair_oldce.log" 2>&1; stamp "bench pair+old ce rc=$?"
D9D_GEMM_PAIR=1 timeout 120 python benchmarks/profile_step.py --out "$out/step_profile_pair.json" > "$out/profile.log" 2>&1; stamp "profile rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel|gemm_fp8_kernel|flash_" -c 12 -f -o "$out/prof_r2" python benchmarks/ncu_targets_r2.py > "$out/ncu.log" 2>&1; stamp "ncu rc=$?"
tail -n 3 "$out/t_ce.log"
grep -h -o '"ms_per_step": [0-9.]*' "$out"/bench_*.log
ls -la "$out"
stamp done
