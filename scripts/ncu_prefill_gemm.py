"""One profiled launch of each prefill projection (OPT-30B, one minibatch: M = 8192) for `ncu --profile-from-start off`:
warm-up launches run outside the profiled range, weights rotate so the profiled launch starts with a cold L2.
  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r1_prof_gemm_prefill_pair \
      python scripts/ncu_prefill_gemm.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lia_b200  # noqa: F401
from lia_b200 import ops

dev = "cuda"
h, f, M = 7168, 28672, 8192
for label, N, K, epi in [("qkv", 3 * h, h, 0), ("out", h, h, 2), ("fc1", f, h, 1), ("fc2", h, f, 2)]:
    ws_ = [torch.randn(N, K, device=dev).to(torch.bfloat16) * 0.02 for _ in range(2)]
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    bias = torch.randn(N, device=dev).to(torch.bfloat16)
    res = torch.randn(M, N, device=dev).to(torch.bfloat16) if epi == 2 else None
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for i in range(2):
        ops.gemm(a, ws_[0], bias, out=out, epilogue=epi, residual=res)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    ops.gemm(a, ws_[1], bias, out=out, epilogue=epi, residual=res)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled", label, flush=True)
    del ws_, a, out, res
    torch.cuda.empty_cache()
