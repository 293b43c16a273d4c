#!/bin/bash
# 8-GPU call (charged 8x box time: keep it under ~6 min): headline workload at TP8, BASELINE.json config 4 (OPT-66B,
# B=64, 512 in / 64 out) at TP8 and config 5b (OPT-175B dummy weights, B=64, 256 in / 32 out, TP8 fully resident).
#   gpurun --gpus 8 --timeout 480 -- bash scripts/gpu_tp8.sh
R=${1:-r2}
N=${2:-8}
mkdir -p gpurun_out
nvidia-smi -L | head -8
run() {  # name, timeout, bench args...
  local name=$1 to=$2; shift 2
  timeout $to python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
    bench.py --gpus $N "$@" > gpurun_out/${R}_${name}_tp${N}.json 2> gpurun_out/${name}_tp${N}.err
  echo "$name exit $?"; cat gpurun_out/${R}_${name}_tp${N}.json; tail -2 gpurun_out/${name}_tp${N}.err
}
run bench_c2 120 --steps 3 --warmup 3
run bench_c4 150 --steps 2 --warmup 3 --model opt-66b --input-tokens 512 --max-new-tokens 64 --num-minibatch 2
run bench_c5b 200 --steps 2 --warmup 3 --model opt-175b --weights dummy --num-minibatch 2
