#!/bin/bash
# decode program (one persistent kernel per decode step): parity against the kernel-per-operation path, then the headline
# bench with and without it IN THE SAME CALL (box-to-box clock differences are larger than the effect being measured).
R=${1:-r2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_decode_program.py -m gpu -x -q > gpurun_out/${R}_pytest_program.log 2>&1
echo "pytest exit $?" >> gpurun_out/${R}_pytest_program.log; tail -30 gpurun_out/${R}_pytest_program.log
if grep -q "pytest exit 0" gpurun_out/${R}_pytest_program.log; then
  for P in 1 0; do
    LIA_DECODE_PROGRAM=$P timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/${R}_bench_program$P.json 2> gpurun_out/bench_program$P.err
    echo "bench program=$P exit $?"; tail -3 gpurun_out/bench_program$P.err
    python -c "
import json,sys
d=json.load(open('gpurun_out/${R}_bench_program$P.json'))
print('program=$P', {k:round(d[k],2) for k in ('value','prefill_ms','decode_ms_per_step')}, 'decode frac', round(d['roofline_decode']['frac'],3), 'launches', d['gpu_launches'], d['clocks'])"
  done
fi
