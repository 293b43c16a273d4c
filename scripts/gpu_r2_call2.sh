#!/bin/bash
# Round-2 GPU call 2 (1 GPU): the new bench.py (parity of the timed result, --config presets, honest CPU arm) on c2 and c1,
# then the ncu evidence for the kernels that run: launch list of the bench command, --set full of the CTA-pair prefill
# GEMM and of the decode GEMMs.
R=${1:-r2}
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[+$(( $(date +%s) - T0 ))s] $*"; }
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/${R}_bench_c2_n1.json 2> gpurun_out/bench_c2.err
echo "bench c2 exit $?"; cat gpurun_out/${R}_bench_c2_n1.json; tail -5 gpurun_out/bench_c2.err; el "c2 done"
timeout 200 python bench.py --config c1 --steps 5 --warmup 3 > gpurun_out/${R}_bench_c1_n1.json 2> gpurun_out/bench_c1.err
echo "bench c1 exit $?"; cat gpurun_out/${R}_bench_c1_n1.json; tail -5 gpurun_out/bench_c1.err; el "c1 done"
timeout 200 python bench.py --config c1 --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_c1_reference.json 2> gpurun_out/bench_c1_ref.err
echo "ref c1 exit $?"; cat gpurun_out/${R}_bench_c1_reference.json; tail -3 gpurun_out/bench_c1_ref.err; el "ref c1 done"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_c2_reference.json 2> gpurun_out/bench_c2_ref.err
echo "ref c2 exit $?"; cat gpurun_out/${R}_bench_c2_reference.json; tail -3 gpurun_out/bench_c2_ref.err; el "ref c2 done"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/${R}_launches_bench_l8.csv \
  python bench.py --layers 8 --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_launches.log 2>&1
echo "ncu launch list exit $?"; el "launch list done"
timeout 150 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/${R}_prof_gemm_prefill_pair \
  python scripts/ncu_prefill_gemm.py > gpurun_out/ncu_prefill_pair.log 2>&1
echo "ncu prefill exit $?"; el "ncu prefill done"
timeout 180 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/${R}_prof_gemm_decode \
  python scripts/ncu_decode_gemm.py > gpurun_out/ncu_decode.log 2>&1
echo "ncu decode exit $?"; el "ncu decode done"
gzip -f gpurun_out/${R}_launches_bench_l8.csv
ls -la gpurun_out | tail -20
