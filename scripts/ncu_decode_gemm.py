"""One profiled launch of each decode projection (OPT-30B, B=64) for `ncu --profile-from-start off`:
warm-up launches run outside the profiled range, weights rotate so the profiled launch misses L2.
  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r1_prof_gemm_decode \
      python scripts/ncu_decode_gemm.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lia_b200  # noqa: F401
from lia_b200 import ops

dev = "cuda"
h, f, B, V = 7168, 28672, 64, 50272
for label, N, K, epi in [("qkv", 3 * h, h, 0), ("out", h, h, 2), ("fc1", f, h, 1), ("fc2", h, f, 2), ("lm_head", V, h, 0)]:
    ws_ = [torch.randn(N, K, device=dev).to(torch.bfloat16) * 0.02 for _ in range(3)]
    a = torch.randn(B, K, device=dev).to(torch.bfloat16)
    bias = torch.randn(N, device=dev).to(torch.bfloat16)
    res = torch.randn(B, N, device=dev).to(torch.bfloat16) if epi == 2 else None
    out = torch.empty(B, N, device=dev, dtype=torch.bfloat16)
    wsp = ops.GemmWorkspace(ops.GemmWorkspace.bytes_for([(B, N, K)]), dev)
    for i in range(3):
        ops.gemm(a, ws_[i % 2], bias, out=out, epilogue=epi, residual=res, workspace=wsp)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    ops.gemm(a, ws_[2], bias, out=out, epilogue=epi, residual=res, workspace=wsp)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled", label, flush=True)
    del ws_
    torch.cuda.empty_cache()
