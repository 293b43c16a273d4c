#!/bin/bash
# BASELINE.json config 3: OPT-30B, batch 512, 256 in / 32 out, gpu-percentage 10 (44 of 48 layers streamed from
# pinned host memory) -- K/V (203 GB) exceed HBM, so the last layers' K/V spill to pinned host memory.
mkdir -p gpurun_out
grep -E "MemTotal|MemAvailable" /proc/meminfo
timeout ${1:-420} python run.py -m opt-30b --batch-size 512 --input-tokens 256 --max-new-tokens 32 --gpu-percentage 10 \
  --num-minibatch 4 --prefill-policy 0 --decoding-policy 0 --pin-weight --num-iter 2 --num-warmup 1 --token-latency --greedy \
  > gpurun_out/config3_run.log 2>&1
echo "config3 exit $?" >> gpurun_out/config3_run.log
tail -14 gpurun_out/config3_run.log
