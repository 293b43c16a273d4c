"""How much of the prefill is attention at the reference's long-prompt settings (its own scripts sweep --input-tokens over
32 / 256 / 1792 / 2016: examples/cpu/inference/python/llm/scripts/lia_offline.sh)?  Times, with CUDA events, one layer's four
projection GEMMs and its causal attention at OPT-30B dims for several prompt lengths, and prints attention's share and its
TFLOP/s.  Decides whether a tcgen05 flash-attention prefill kernel is worth writing (not built).
  python scripts/long_prompt_probe.py [rows=8192]      rows = tokens per minibatch (B*S is held near this)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import lia_b200  # noqa: E402,F401
from lia_b200 import ops  # noqa: E402

dev, BF16 = "cuda", torch.bfloat16
h, f, H, d = 7168, 28672, 56, 128
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 8192


def timeit(fn, n=5, warm=2):
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


wq = (torch.randn(3 * h, h, device=dev) * 0.02).to(BF16)
wo = (torch.randn(h, h, device=dev) * 0.02).to(BF16)
w1 = (torch.randn(f, h, device=dev) * 0.02).to(BF16)
w2 = (torch.randn(h, f, device=dev) * 0.02).to(BF16)
bq, bo, b1 = (torch.zeros(n, device=dev, dtype=BF16) for n in (3 * h, h, f))
for S in (256, 512, 1024, 1792, 2016):
    B = max(1, rows // S)
    M = B * S
    x = torch.randn(M, h, device=dev).to(BF16)
    q = torch.empty(M, h, device=dev, dtype=BF16)
    ctx = torch.empty(M, h, device=dev, dtype=BF16)
    ffn = torch.empty(M, f, device=dev, dtype=BF16)
    out = torch.empty(M, h, device=dev, dtype=BF16)
    kc = torch.zeros(S, B, H, d, device=dev, dtype=BF16)
    vc = torch.zeros_like(kc)
    args = ops.qkv_args(q, kc, vc, S, 0, 0, d ** -0.5)
    t_qkv = timeit(lambda: ops.gemm(x, wq, bq, epilogue=ops.EPI_QKV, qkv=args))
    t_att = timeit(lambda: ops.attn_prefill(q, kc, vc, B, S, 0, out=ctx))
    t_o = timeit(lambda: ops.gemm(ctx, wo, bo, out=out, epilogue=ops.EPI_BIAS_RESIDUAL, residual=x))
    t_1 = timeit(lambda: ops.gemm(x, w1, b1, out=ffn, epilogue=ops.EPI_BIAS_RELU))
    t_2 = timeit(lambda: ops.gemm(ffn, w2, bo, out=out, epilogue=ops.EPI_BIAS_RESIDUAL, residual=x))
    gemm = t_qkv + t_o + t_1 + t_2
    fl_att = 2.0 * S * S * h * B                      # causal Q.K^T + P.V  (4*S^2*h*B / 2)
    print(f"S={S:5d} B={B:3d}: GEMMs {gemm:8.3f} ms ({24.0 * h * h * M / gemm / 1e9:7.1f} TFLOP/s)   attention {t_att:8.3f} ms "
          f"({fl_att / t_att / 1e9:6.1f} TFLOP/s algorithmic)   attention share of the layer {t_att / (gemm + t_att) * 100:5.1f} %", flush=True)
    del x, q, ctx, ffn, out, kc, vc
    torch.cuda.empty_cache()
