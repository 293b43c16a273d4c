#!/bin/bash
# 2-GPU call: the multi-GPU exchange test, then the headline bench at TP2 and BASELINE.json config 4 (OPT-66B, B=64,
# 512 in / 64 out) at TP2.
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python -m pytest tests/test_gpu_tp.py -x -q -m gpu > gpurun_out/pytest_tp2.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_tp2.log
tail -5 gpurun_out/pytest_tp2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
  bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 exit $?"; cat gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 \
  bench.py --gpus 2 --steps 2 --warmup 3 --model opt-66b --input-tokens 512 --max-new-tokens 64 --num-minibatch 2 \
  > gpurun_out/bench_c4_tp2.json 2> gpurun_out/bench_c4_tp2.err
echo "bench c4 exit $?"; cat gpurun_out/bench_c4_tp2.json; tail -3 gpurun_out/bench_c4_tp2.err
