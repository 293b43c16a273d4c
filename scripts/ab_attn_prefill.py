"""A/B of the two prefill attention kernels on a GPU: the tcgen05 kernel (default, attn_prefill_sm100.cu) against the
first-generation mma.sync kernel (LIA_ATTN_PREFILL_TC=0).  Both form s = bf16(q.k), p = bf16(softmax) and ctx = bf16(p.v) at
the reference's rounding points; they differ in fp32 summation order and in exp/division arithmetic (ex2.approx and a
multiply by 1/l vs expf and a true division), so outputs agree to a few bf16 ulps of the row maximum, not bit for bit.
Prints per shape: elements that differ, the largest difference in bf16 ulps of the row max, and both timings.
  python scripts/ab_attn_prefill.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import lia_b200  # noqa: E402,F401
from lia_b200 import ops  # noqa: E402

dev = "cuda"
BF16 = torch.bfloat16


def run(tc, q, kc, vc, B, S):
    os.environ["LIA_ATTN_PREFILL_TC"] = tc
    return ops.attn_prefill(q, kc, vc, B, S, 0)


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


worst = 0.0
for (B, H, S, d) in [(2, 2, 17, 128), (1, 1, 1, 64), (2, 3, 64, 128), (3, 5, 100, 64), (5, 2, 129, 64), (4, 7, 200, 128), (2, 4, 255, 128),
                     (8, 32, 256, 64), (2, 8, 300, 128), (2, 4, 1000, 64), (32, 56, 256, 128), (16, 56, 512, 128), (4, 56, 2016, 128)]:
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + S)
    q = (torch.randn(B * S, H * d, generator=g) * 0.3).to(BF16).to(dev)
    kc = (torch.randn(S + 3, B, H, d, generator=g)).to(BF16).to(dev)
    vc = (torch.randn(S + 3, B, H, d, generator=g)).to(BF16).to(dev)
    y0 = run("0", q, kc, vc, B, S)
    y1 = run("1", q, kc, vc, B, S)
    torch.cuda.synchronize()
    scale = y0.float().view(B * S, H, d).abs().amax(-1, keepdim=True).clamp_min(1e-30)
    ulp = 2.0 ** (torch.floor(torch.log2(scale)) - 7)
    du = ((y1.float() - y0.float()).view(B * S, H, d).abs() / ulp)
    nd = int((y0.view(torch.int16) != y1.view(torch.int16)).sum())
    worst = max(worst, du.max().item())
    line = (f"B={B:3d} H={H:3d} S={S:4d} d={d:4d}: {nd:9d} of {y0.numel():9d} differ, max {du.max().item():5.2f} ulp of the row max, "
            f"nan={bool(torch.isnan(y1.float()).any())}")
    if B * H * S >= 100000:
        t0 = timeit(lambda: run("0", q, kc, vc, B, S))
        t1 = timeit(lambda: run("1", q, kc, vc, B, S))
        fl = 2.0 * S * S * H * d * B
        line += f";  mma.sync {t0 * 1e3:8.1f} us ({fl / t0 / 1e9:6.1f} TFLOP/s)   tcgen05 {t1 * 1e3:8.1f} us ({fl / t1 / 1e9:6.1f} TFLOP/s, x{t0 / t1:.2f})"
    print(line, flush=True)
print("AB_ATTN_PREFILL worst", worst, flush=True)
