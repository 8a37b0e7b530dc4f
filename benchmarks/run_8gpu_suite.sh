#!/bin/bash
# One 8-GPU session: correctness over the reference's meshes, NVLink kernels at 8 ranks, EP / reference-example / 30B-A3B
# bench layouts.  Every step has its own timeout (a hung collective must not burn the box); logs go to gpurun_out/.
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
echo "== validate meshes (8 GPUs)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29731 \
  benchmarks/validate_parallelism_gpu.py --skip-attention 2>gpurun_out/validate8.err | tee gpurun_out/validate_parallelism_w8.log | tail -20
echo "== nvlink tests at 8 ranks"
timeout 300 python -m pytest tests/test_nvlink_gpu.py -q -x -k "more_gpus and 0-8 or matches_local_experts and 4-8 or matches_local_experts and 0-8" 2>&1 | tail -4 | tee gpurun_out/nvlink8.log
for L in ep example; do
  echo "== bench --layout $L"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29741 \
    bench.py --gpus 8 --steps 5 --warmup 3 --layout $L 2>gpurun_out/bench8_$L.err | tail -1 | tee gpurun_out/bench8_$L.json | cut -c1-700
done
echo "== bench 30B-A3B shape, EP8"
timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29751 \
  bench.py --gpus 8 --steps 3 --warmup 2 --layout ep --model 30b-a3b --checkpointing --ep-capacity-factor 1.5 2>gpurun_out/bench8_30b.err | tail -1 | tee gpurun_out/bench8_30b.json | cut -c1-900
tail -3 gpurun_out/bench8_30b.err | cut -c1-400
