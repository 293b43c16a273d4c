#!/bin/bash
# tcgen05 prefill attention: A/B against the mma.sync kernel, then the kernel parity tests.  Each step under its own timeout
# (a wrong barrier protocol hangs rather than fails).
mkdir -p gpurun_out
timeout 120 python scripts/ab_attn_prefill.py > gpurun_out/r2_ab_attn_tc.log 2>&1; echo "ab exit $?"; tail -20 gpurun_out/r2_ab_attn_tc.log
timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attn_prefill" > gpurun_out/r2_pytest_attn.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/r2_pytest_attn.log
