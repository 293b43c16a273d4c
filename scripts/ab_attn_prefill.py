"""A/B of the two prefill attention kernels on a GPU: the default two-pass kernel (scores recomputed in pass 2) against the
opt-in variant that keeps the bf16-rounded scores of pass 1 in registers (LIA_ATTN_PREFILL_KEEP=1, S <= 256).
The variant consumes exactly the values the default recomputes, so the outputs must be BIT-IDENTICAL; the script checks
that on every shape, then times both at the headline shape (OPT-30B, one minibatch: B=32, H=56, S=256, d=128).
Also reports the opt-in fast-math softmax (LIA_ATTN_FASTMATH=1: __expf and multiply by 1/l -- same bf16 rounding points,
a few fp32 ulps before them): how many outputs differ from the default, by how many bf16 ulps, and the speed-up, including
at a long prompt (S=2016), where attention is a large share of the prefill.
  python scripts/ab_attn_prefill.py        -> prints one line per shape and a final verdict line"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import lia_b200  # noqa: E402,F401
from lia_b200 import ops  # noqa: E402

dev = "cuda"
BF16 = torch.bfloat16


def run(flag, q, kc, vc, B, S, fast="0"):
    os.environ["LIA_ATTN_PREFILL_KEEP"] = flag
    os.environ["LIA_ATTN_FASTMATH"] = fast
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


ok = True
for (B, H, S, d) in [(32, 56, 256, 128), (4, 56, 2016, 128), (8, 32, 256, 64), (3, 5, 100, 64), (2, 3, 64, 128), (2, 2, 17, 128), (1, 1, 1, 64),
                     (4, 7, 200, 128), (5, 2, 129, 64), (2, 4, 255, 128)]:
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + S)
    q = (torch.randn(B * S, H * d, generator=g) * 0.3).to(BF16).to(dev)
    kc = (torch.randn(S + 3, B, H, d, generator=g)).to(BF16).to(dev)
    vc = (torch.randn(S + 3, B, H, d, generator=g)).to(BF16).to(dev)
    y0 = run("0", q, kc, vc, B, S)
    y1 = run("1", q, kc, vc, B, S)
    torch.cuda.synchronize()
    nd = int((y0.view(torch.int16) != y1.view(torch.int16)).sum())
    ok = ok and nd == 0
    line = f"B={B:3d} H={H:3d} S={S:4d} d={d:4d}: {nd} of {y0.numel()} elements differ"
    # LIA_ATTN_FASTMATH=1 (__expf, multiply by 1/l): not bit-identical by design -- report how far, in bf16 ulps of the
    # output, and how often; the rounding points are the same, so differences should be rare single-ulp flips
    yf = run("0", q, kc, vc, B, S, fast="1")
    yfk = run("1", q, kc, vc, B, S, fast="1")
    torch.cuda.synchronize()
    ulp = 2.0 ** (torch.floor(torch.log2(y0.float().abs().clamp_min(1e-30))) - 7)
    du = ((yf.float() - y0.float()).abs() / ulp)
    line += f";  fastmath: {int((du > 0).sum())} differ, max {du.max().item():.1f} ulp, keep+fast == fast: {bool(torch.equal(yf, yfk))}"
    if B * H * S >= 100000:
        t0 = timeit(lambda: run("0", q, kc, vc, B, S))
        t1 = timeit(lambda: run("1", q, kc, vc, B, S))
        t2 = timeit(lambda: run("0", q, kc, vc, B, S, fast="1"))
        t3 = timeit(lambda: run("1", q, kc, vc, B, S, fast="1"))
        line += (f";  two-pass {t0 * 1e3:8.1f} us   keep-scores {t1 * 1e3:8.1f} us (x{t0 / t1:.2f})   fastmath {t2 * 1e3:8.1f} us "
                 f"(x{t0 / t2:.2f})   keep+fastmath {t3 * 1e3:8.1f} us (x{t0 / t3:.2f})")
    print(line, flush=True)
print("AB_ATTN_PREFILL", "BIT-IDENTICAL" if ok else "MISMATCH", flush=True)
