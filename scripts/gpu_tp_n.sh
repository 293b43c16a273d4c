#!/bin/bash
# N-GPU call: the multi-GPU exchange test (N = 2 only), then the headline bench at TP-N with parity of the timed result.
#   gpurun --gpus N --timeout 600 -- bash scripts/gpu_tp_n.sh N [tag]
N=${1:-2}
R=${2:-r2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
if [ "$N" = "2" ]; then
  timeout 400 python -m pytest tests/test_gpu_tp.py -x -q -m gpu -k multi_gpu > gpurun_out/${R}_pytest_tp2.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${R}_pytest_tp2.log; tail -8 gpurun_out/${R}_pytest_tp2.log
fi
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
  bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${R}_bench_c2_tp$N.json 2> gpurun_out/bench_tp$N.err
echo "bench tp$N exit $?"; cat gpurun_out/${R}_bench_c2_tp$N.json; tail -4 gpurun_out/bench_tp$N.err
