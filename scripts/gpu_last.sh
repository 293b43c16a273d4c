#!/bin/bash
# Last GPU minutes of round 1: headline bench with the CTA-pair prefill GEMM as the default, then one ncu --set full
# capture of its four prefill projections.
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[+$(( $(date +%s) - T0 ))s] $*"; }
timeout 120 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1_pair.json 2> gpurun_out/bench_n1_pair.err
echo "bench exit $?"; cat gpurun_out/bench_n1_pair.json; tail -3 gpurun_out/bench_n1_pair.err
el "bench done"
timeout 70 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r1_prof_gemm_prefill_pair \
  python scripts/ncu_prefill_gemm.py > gpurun_out/ncu_prefill_pair.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_prefill_pair.log
el "ncu done"
