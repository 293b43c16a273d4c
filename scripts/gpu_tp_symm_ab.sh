#!/bin/bash
# A/B in ONE call at TP-N: symmetric-memory arena (NVLS prefill exchange) vs CUDA-IPC arena (peer stores)
N=${1:-4}
mkdir -p gpurun_out
for SY in 1 0; do
  LIA_TP_SYMM=$SY timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
    bench.py --gpus $N --steps 3 --warmup 3 --no-parity > gpurun_out/r2_bench_c2_tp${N}_symm$SY.json 2> gpurun_out/bench_symm.err
  python -c "
import json
d=json.load(open('gpurun_out/r2_bench_c2_tp${N}_symm$SY.json'))
r=d['roofline']
print('symm=$SY', {k:round(d[k],2) for k in ('value','prefill_ms','decode_ms_per_step')}, 'col', round(r['column_parallel_tflops']), 'row+AR', round(r['row_parallel_fused_allreduce_tflops']), d['clocks'])"
done
