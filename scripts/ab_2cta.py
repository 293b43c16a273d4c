"""A/B of the prefill GEMM: one-CTA 128x256 tiles vs CTA pairs (LIA_GEMM_2CTA=1), OPT-30B prefill shapes, M = 8192.
CUDA events, weights rotate over buffers > L2.  Usage: python scripts/ab_2cta.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lia_b200  # noqa: F401
from lia_b200 import ops

dev = "cuda"
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


for label, N, K, epi in [("qkv", 3 * h, h, 0), ("out", h, h, 2), ("fc1", f, h, 1), ("fc2", h, f, 2)]:
    ws_ = [torch.randn(N, K, device=dev).to(torch.bfloat16) * 0.02 for _ in range(2)]
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    bias = torch.randn(N, device=dev).to(torch.bfloat16)
    res = torch.randn(M, N, device=dev).to(torch.bfloat16) if epi == 2 else None
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    outs = {}
    for flag, bn224 in (("0", "0"), ("1", "0")):
        os.environ["LIA_GEMM_2CTA"] = flag
        ms = timeit(lambda i: ops.gemm(a, ws_[i % 2], bias, out=out, epilogue=epi, residual=res))
        outs[(flag, bn224)] = ops.gemm(a, ws_[0], bias, epilogue=epi, residual=res)
        torch.cuda.synchronize()
        same = "" if (flag, bn224) == ("0", "0") else f"  bit-identical to one-CTA: {bool(torch.equal(outs[(flag, bn224)], outs[('0', '0')]))}"
        print(f"{label:4s} M={M} N={N:6d} K={K:6d} 2cta={flag} bn224={bn224}: {ms * 1e3:9.1f} us  {2.0 * M * N * K / ms / 1e9:8.1f} TFLOP/s{same}", flush=True)
    outs.clear()
    del ws_, a, out, res
    torch.cuda.empty_cache()
