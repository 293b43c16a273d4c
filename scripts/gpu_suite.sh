#!/bin/bash
# whole GPU parity suite + one headline bench line (1 GPU).  Usage: gpurun --timeout 900 -- bash scripts/gpu_suite.sh [tag]
R=${1:-r2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/${R}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${R}_pytest_gpu.log; tail -12 gpurun_out/${R}_pytest_gpu.log
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_c2_n1.json 2> gpurun_out/bench_c2.err
echo "bench c2 exit $?"; cut -c1-1200 gpurun_out/${R}_bench_c2_n1.json; tail -5 gpurun_out/bench_c2.err
R=$R python -c "
import json,os
d=json.load(open('gpurun_out/%s_bench_c2_n1.json' % os.environ['R']))
print({k:round(d[k],2) for k in ('value','prefill_ms','decode_ms_per_step')}, round(d['roofline']['frac'],3), round(d['roofline_decode']['frac'],3), d.get('parity'))"
