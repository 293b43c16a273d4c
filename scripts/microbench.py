"""Per-kernel CUDA-event timings at the bench shapes (OPT-30B, B=64): decode GEMMs (HBM-bound),
decode attention, LayerNorm, prefill GEMMs.  Weights rotate over several buffers so that no
launch finds its operands in L2."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lia_b200
from lia_b200 import ops

dev = "cuda"
h, f, H, d, B, V = 7168, 28672, 56, 128, 64, 50272
HBM = 6453.1

def timeit(fn, n=20, warm=3):
    for i in range(warm): fn(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(n): fn(i)
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

def bench_gemm(M, N, K, epi, nbuf=4, label=""):
    ws_ = [torch.randn(N, K, device=dev).to(torch.bfloat16) * 0.02 for _ in range(nbuf)]
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    bias = torch.randn(N, device=dev).to(torch.bfloat16)
    res = torch.randn(M, N, device=dev).to(torch.bfloat16) if epi == 2 else None
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    wsp = ops.GemmWorkspace(ops.GemmWorkspace.bytes_for([(M, N, K)]), dev)
    ms = timeit(lambda i: ops.gemm(a, ws_[i % nbuf], bias, out=out, epilogue=epi, residual=res, workspace=wsp))
    byts = 2.0 * N * K + 2.0 * M * K + 2.0 * M * N
    fl = 2.0 * M * N * K
    print(f"gemm {label:8s} M={M:6d} N={N:6d} K={K:6d} epi={epi} streamk_env={os.environ.get('LIA_STREAMK','-')}: {ms*1e3:9.1f} us  {byts/ms/1e6:8.1f} GB/s ({byts/ms/1e6/HBM*100:5.1f}% HBM)  {fl/ms/1e9:8.1f} TFLOP/s", flush=True)
    return ms

def bench_attn(T, nbuf=3):
    kcs = [torch.randn(T + 8, B, H, d, device=dev).to(torch.bfloat16) for _ in range(nbuf)]
    vcs = [torch.randn(T + 8, B, H, d, device=dev).to(torch.bfloat16) for _ in range(nbuf)]
    q = torch.randn(B, H * d, device=dev).to(torch.bfloat16)
    out = torch.empty(B, H * d, device=dev, dtype=torch.bfloat16)
    ws = ops.attn_decode_workspace(B, H, d, dev)
    ms = timeit(lambda i: ops.attn_decode(q, kcs[i % nbuf], vcs[i % nbuf], B, T, 0, out=out, workspace=ws))
    byts = 4.0 * B * T * H * d
    print(f"attn_decode T={T}: {ms*1e3:9.1f} us  {byts/ms/1e6:8.1f} GB/s ({byts/ms/1e6/HBM*100:5.1f}% HBM)", flush=True)

def bench_ln(rows):
    x = torch.randn(rows, h, device=dev).to(torch.bfloat16)
    w = torch.ones(h, device=dev, dtype=torch.bfloat16); b = torch.zeros(h, device=dev, dtype=torch.bfloat16)
    y = torch.empty_like(x)
    ms = timeit(lambda i: ops.layernorm(x, w, b, out=y))
    print(f"layernorm rows={rows}: {ms*1e3:9.1f} us  {4.0*rows*h/ms/1e6:8.1f} GB/s", flush=True)

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "decode"):
    bench_gemm(B, 3 * h, h, 0, label="qkv")
    bench_gemm(B, h, h, 2, label="out")
    bench_gemm(B, f, h, 1, label="fc1")
    bench_gemm(B, h, f, 2, label="fc2")
    bench_gemm(B, V, h, 0, label="lm_head")
    for T in (257, 272, 288): bench_attn(T)
    bench_ln(64)
if which in ("all", "prefill"):
    M = 8192
    bench_gemm(M, 3 * h, h, 0, nbuf=2, label="qkv")
    bench_gemm(M, h, h, 2, nbuf=2, label="out")
    bench_gemm(M, f, h, 1, nbuf=2, label="fc1")
    bench_gemm(M, h, f, 2, nbuf=2, label="fc2")
    bench_ln(M)
