#!/bin/bash
# 1-GPU call: teacher-forced token parity at the FULL headline config, headline bench with per-layer parity, BASELINE.json configs 3 and 5a.
R=${1:-r2}
mkdir -p gpurun_out
grep -E "MemTotal|MemAvailable" /proc/meminfo
timeout 400 python scripts/token_parity_report.py opt-30b 0 64 256 32 > gpurun_out/${R}_token_parity_30b_full.json 2> gpurun_out/token_parity_full.err
echo "token parity exit $?"; python -c "
import json; d=json.load(open('gpurun_out/${R}_token_parity_30b_full.json')); d['flip_list']=d['flip_list'][:6]; print(d)"; tail -2 gpurun_out/token_parity_full.err
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/${R}_bench_c2_n1_final.json 2> gpurun_out/bench_c2.err
echo "bench c2 exit $?"; python -c "
import json; d=json.load(open('gpurun_out/${R}_bench_c2_n1_final.json')); print({k:d[k] for k in ('value','prefill_ms','decode_ms_per_step')}, d['parity'], d.get('cpu_baseline',{}).get('value'))"
timeout 500 python bench.py --config c3 --steps 1 --quick --no-cpu-baseline --no-parity > gpurun_out/${R}_bench_c3_n1.json 2> gpurun_out/bench_c3.err
echo "bench c3 exit $?"; cat gpurun_out/${R}_bench_c3_n1.json | cut -c1-2500; tail -3 gpurun_out/bench_c3.err
LIA_HOST_LAYER_POOL=24 timeout 700 python bench.py --config c5a --steps 1 --quick --no-cpu-baseline > gpurun_out/${R}_bench_c5a_n1.json 2> gpurun_out/bench_c5a.err
echo "bench c5a exit $?"; cat gpurun_out/${R}_bench_c5a_n1.json | cut -c1-2500; tail -3 gpurun_out/bench_c5a.err
