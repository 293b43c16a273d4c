#!/bin/bash
# One short gpurun call (GPU budget nearly spent): parity suite, then the opt-in CTA-pair GEMM (test + A/B), then an
# ncu --set full capture of the decode projections, then the headline bench if time is left.  Every step has its own timeout.
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[+$(( $(date +%s) - T0 ))s] $*"; }
timeout 420 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
el "pytest done"; tail -22 gpurun_out/pytest_gpu.log
LIA_TEST_2CTA=1 timeout 150 python -m pytest tests/test_gpu_gemm_2cta.py -m gpu -x -q -v --timeout 45 > gpurun_out/pytest_2cta.log 2>&1
RC2=$?
echo "pytest 2cta exit $RC2" >> gpurun_out/pytest_2cta.log
el "2cta test done"; tail -25 gpurun_out/pytest_2cta.log
if [ "$RC2" = "0" ]; then
  timeout 120 python scripts/ab_2cta.py > gpurun_out/ab_2cta.log 2>&1
  echo "ab exit $?" >> gpurun_out/ab_2cta.log
  el "ab done"; cat gpurun_out/ab_2cta.log | tail -12
fi
timeout 150 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r1_prof_gemm_decode \
  python scripts/ncu_decode_gemm.py > gpurun_out/ncu_decode.log 2>&1
echo "ncu exit $?" >> gpurun_out/ncu_decode.log
timeout 60 ncu -i gpurun_out/r1_prof_gemm_decode.ncu-rep --page raw --csv > gpurun_out/r1_prof_gemm_decode.raw.csv 2>> gpurun_out/ncu_decode.log
el "ncu done"; tail -4 gpurun_out/ncu_decode.log
timeout 280 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench exit $?"; cat gpurun_out/bench_n1.json
el "bench done"
