#!/bin/bash
# A/B of the prefill exchange at TP-N in ONE call: NVLS (in-switch reduce + multicast store) vs peer stores
N=${1:-8}
mkdir -p gpurun_out
for NV in 1 0; do
  LIA_TP_NVLS=$NV timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
    bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_c2_tp${N}_nvls$NV.json 2> gpurun_out/bench_nvls.err
  echo "nvls=$NV exit $?"; tail -2 gpurun_out/bench_nvls.err | cut -c1-300
  python -c "
import json
d=json.load(open('gpurun_out/r2_bench_c2_tp${N}_nvls$NV.json'))
r=d['roofline']
print('nvls=$NV', {k:round(d[k],2) for k in ('value','prefill_ms','decode_ms_per_step')}, 'col', round(r['column_parallel_tflops']), 'row+AR', round(r['row_parallel_fused_allreduce_tflops']), 'row ms/prefill', round(r['row_parallel_fused_allreduce_ms_per_prefill'],1), d['parity'].get('prefill_hidden_rel_err'), d['parity'].get('tokens_equal_frac'), d['parity'].get('worst_first_divergence_margin_bf16_ulps'))"
done
