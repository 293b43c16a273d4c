#!/bin/bash
# A/B of the L2 prefetch-next hints (LIA_L2_PREFETCH_MB) on the headline decode step, all settings in ONE call (same box, same clocks)
mkdir -p gpurun_out
for MB in ${@:-0 16 32 48 64 96}; do
  LIA_L2_PREFETCH_MB=$MB timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2_bench_pf$MB.json 2> gpurun_out/bench_pf.err
  python -c "
import json
d=json.load(open('gpurun_out/r2_bench_pf$MB.json'))
print('prefetch_mb=$MB', {k:round(d[k],2) for k in ('value','prefill_ms','decode_ms_per_step')}, 'decode frac', round(d['roofline_decode']['frac'],3), d['clocks'])"
done
