#!/bin/bash
# final 1-GPU call of a round: whole GPU suite, headline bench line, decode-exchange microbench on one GPU (self-loop)
R=${1:-r2}
bash scripts/gpu_suite.sh $R
timeout 100 python scripts/tp_sim_trace.py 8 2>&1 | grep -E "^out_proj|^fc2|^qkv|^fc1" | tee gpurun_out/${R}_tp_sim_trace.log
LIA_TP_SIM_WORLD=8 timeout 200 python scripts/tp_decode_probe.py 16 2>/dev/null | grep PROBE | tee -a gpurun_out/${R}_tp_sim_trace.log
