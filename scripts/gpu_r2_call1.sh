#!/bin/bash
# Round-2 GPU call 1 (1 GPU): whole parity suite WITHOUT -x (every failure reports), the opt-in variants, a headline bench
# line, the queued A/B scripts and the teacher-forced token-parity statistics.  Results in gpurun_out/.
#   gpurun --timeout 1200 -- bash scripts/gpu_r2_call1.sh
R=${1:-r2}
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[+$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
timeout 600 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${R}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${R}_pytest_gpu.log
el "pytest done"; tail -25 gpurun_out/${R}_pytest_gpu.log
LIA_TEST_OPTIN=1 timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_gemm_2cta.py -m gpu -q -k "optin or bn224" > gpurun_out/${R}_pytest_optin.log 2>&1
echo "pytest optin exit $?" >> gpurun_out/${R}_pytest_optin.log; tail -6 gpurun_out/${R}_pytest_optin.log
el "optin done"
timeout 240 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_n1_call1.json 2> gpurun_out/bench_n1.err
echo "bench exit $?"; cat gpurun_out/${R}_bench_n1_call1.json; tail -3 gpurun_out/bench_n1.err
el "bench done"
timeout 120 python scripts/ab_2cta.py > gpurun_out/${R}_ab_2cta.log 2>&1; tail -8 gpurun_out/${R}_ab_2cta.log
LIA_GEMM_BN224=1 timeout 120 python scripts/ab_2cta.py > gpurun_out/${R}_ab_2cta_bn224.log 2>&1; tail -8 gpurun_out/${R}_ab_2cta_bn224.log
timeout 120 python scripts/microbench.py decode > gpurun_out/${R}_microbench_decode.log 2>&1; tail -10 gpurun_out/${R}_microbench_decode.log
timeout 120 python scripts/ab_attn_prefill.py > gpurun_out/${R}_ab_attn_prefill.log 2>&1; tail -12 gpurun_out/${R}_ab_attn_prefill.log
timeout 150 python scripts/long_prompt_probe.py > gpurun_out/${R}_long_prompt_probe.log 2>&1; tail -6 gpurun_out/${R}_long_prompt_probe.log
el "microbench done"
timeout 150 python scripts/token_parity_report.py opt-1.3b 3 8 256 32 > gpurun_out/${R}_token_parity_1p3b.json 2> gpurun_out/token_parity.err; tail -c 700 gpurun_out/${R}_token_parity_1p3b.json
timeout 200 python scripts/token_parity_report.py opt-30b 2 8 64 32 > gpurun_out/${R}_token_parity_30b.json 2>> gpurun_out/token_parity.err; tail -c 700 gpurun_out/${R}_token_parity_30b.json
tail -3 gpurun_out/token_parity.err
el "token parity done"
ls -la gpurun_out
