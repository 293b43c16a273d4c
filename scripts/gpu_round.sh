#!/bin/bash
# One gpurun call: GPU parity tests, headline bench, then the config-3 probe (KV spill + weight streaming).
# Usage (from the repo root on the GPU box): bash scripts/gpu_round.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
grep -E "MemTotal|MemAvailable" /proc/meminfo >> gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt
timeout 420 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench exit $?"
cat gpurun_out/bench_n1.json
avail_gb=$(awk '/MemAvailable/ {print int($2/1048576)}' /proc/meminfo)
if [ "$avail_gb" -ge 260 ]; then
  timeout 420 python run.py -m opt-30b --batch-size 512 --input-tokens 256 --max-new-tokens 32 --gpu-percentage 10 \
    --num-minibatch 4 --prefill-policy 0 --decoding-policy 0 --pin-weight --num-iter 2 --num-warmup 1 --token-latency --greedy \
    > gpurun_out/config3_run.log 2>&1
  echo "config3 exit $?" >> gpurun_out/config3_run.log
  tail -12 gpurun_out/config3_run.log
else
  echo "config3 skipped: only ${avail_gb} GB of host memory available" | tee gpurun_out/config3_run.log
fi
