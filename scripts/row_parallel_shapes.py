"""CUDA-event timing of the row-parallel prefill projections at the shapes a TP-N prefill runs (out_proj K = h/N, fc2 K = f/N; plain
GEMM + residual, no exchange): short-K GEMMs are bound by the epilogue, not by the MMAs.  python scripts/row_parallel_shapes.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lia_b200  # noqa: F401
from lia_b200 import ops

dev, BF16 = "cuda", torch.bfloat16
h, f, M = 7168, 28672, 8192


def timeit(fn, n=10, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(n):
        fn(i)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


for world in (1, 2, 4, 8):
    for label, K in (("out_proj", h // world), ("fc2", f // world)):
        ws_ = [torch.randn(h, K, device=dev).to(BF16) * 0.02 for _ in range(2)]
        a = torch.randn(M, K, device=dev).to(BF16)
        bias = torch.randn(h, device=dev).to(BF16)
        res = torch.randn(M, h, device=dev).to(BF16)
        out = torch.empty(M, h, device=dev, dtype=BF16)
        ms = timeit(lambda i: ops.gemm(a, ws_[i % 2], bias, out=out, epilogue=ops.EPI_BIAS_RESIDUAL, residual=res))
        fl = 2.0 * M * h * K
        print(f"TP{world} {label:8s} M={M} N={h} K={K:6d}: {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s", flush=True)
        del ws_, a, res, out
