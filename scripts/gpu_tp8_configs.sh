#!/bin/bash
# 8-GPU call (charged 8x): headline at TP8, then BASELINE.json config 4 (OPT-66B 512/64) and 5b (OPT-175B dummy weights, resident).
R=${1:-r2}
N=${2:-8}
mkdir -p gpurun_out
run() {  # name, timeout, bench args...
  local name=$1 to=$2; shift 2
  timeout $to python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
    bench.py --gpus $N "$@" > gpurun_out/${R}_bench_${name}_tp${N}.json 2> gpurun_out/${name}_tp${N}.err
  echo "$name exit $?"; cat gpurun_out/${R}_bench_${name}_tp${N}.json | cut -c1-3500; tail -2 gpurun_out/${name}_tp${N}.err
}
run c2 150 --steps 3 --warmup 3
run c4 200 --config c4 --steps 2 --warmup 3
if [ "$N" = "8" ]; then run c5b 260 --config c5b --steps 2 --warmup 3; fi
