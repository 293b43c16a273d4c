#!/bin/bash
# BASELINE.json config 5a: OPT-175B dummy weights on ONE B200, non-resident layers streamed from pinned host memory.
# The box has less host RAM (~196 GB available) than the streamed weights (~280 GB), so the streamed layers alias a
# pool of distinct pinned slabs (LIA_HOST_LAYER_POOL): PCIe bytes per step are those of the full model.
mkdir -p gpurun_out
grep -E "MemTotal|MemAvailable" /proc/meminfo
LIA_HOST_LAYER_POOL=${2:-24} timeout ${1:-400} python run.py -m opt-175b --dummy-weights --batch-size 64 --input-tokens 256 \
  --max-new-tokens 32 --gpu-percentage 20 --num-minibatch 2 --prefill-policy 0 --decoding-policy 0 --pin-weight \
  --num-iter 1 --num-warmup 0 --token-latency --greedy > gpurun_out/config5a_run.log 2>&1
echo "config5a exit $?" >> gpurun_out/config5a_run.log
tail -14 gpurun_out/config5a_run.log
