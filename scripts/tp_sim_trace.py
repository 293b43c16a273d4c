"""python scripts/tp_sim_trace.py [WORLD]: ONE GPU, the per-rank decode shapes of a WORLD-way tensor-parallel OPT-30B.
The row-parallel projections run plain (no exchange) and fused with the exchange over a self-loop arena (every "peer" is this
GPU, nobody is waited for: results are garbage, the instruction path and its memory traffic are the real ones), as CUDA graphs
of 32 launches; with the per-CTA timeline (LIA_GEMM_TRACE) of the fused kernel.  Separates what the exchange CODE costs from
what NVLink latency and rank skew cost (scripts/tp_microbench.py measures the sum on real ranks)."""
import ctypes, os, sys
os.environ["LIA_TP_SELF_LOOP"] = "1"
os.environ.setdefault("LIA_TP_NO_WAIT", "1")
os.environ["LIA_GEMM_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import lia_b200
from lia_b200 import _lib, graphs, ops, tp

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
lib = _lib.load()
BF16 = torch.bfloat16
h, f, B = 7168, 28672, 64
NB = 4


def timeit(fn, n, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(n):
        fn(i)
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


def graphed(fn, n=32):
    for i in range(3):        # first launches allocate (trace buffer, function attributes): not under capture
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with graphs.capture(g):
        for i in range(n):
            fn(i)
    return (lambda _i: g.replay()), n


cdll = ctypes.CDLL(_lib.LIB_PATH)
cdll.lia_debug_gemm_trace.restype = ctypes.POINTER(ctypes.c_ulonglong)
for (M, N, K, label) in [(B, h, h // world, "out_proj"), (B, h, f // world, "fc2"), (B, 3 * h // world, h, "qkv (plain only)"), (B, f // world, h, "fc1 (plain only)")]:
    ws_ = [(torch.randn(N, K, device=dev) * 0.02).to(BF16) for _ in range(NB)]
    a = torch.randn(M, K, device=dev).to(BF16)
    bias = torch.randn(N, device=dev).to(BF16)
    res = torch.randn(M, N, device=dev).to(BF16)
    out = torch.empty(M, N, device=dev, dtype=BF16)
    wsp = ops.GemmWorkspace(ops.GemmWorkspace.bytes_for([(M, N, K)]), dev)
    plain = lambda i: ops.gemm(a, ws_[i % NB], bias, out=out, epilogue=ops.EPI_BIAS_RESIDUAL, residual=res, workspace=wsp)
    r = {}
    gfn, gn = graphed(plain)
    r["plain"] = timeit(gfn, n=8) / gn
    if "plain only" not in label:
        arena = tp.PeerArena(0, world, dev, lib.lia_tp_recv_bytes(M, N, K, world), [("pad", 256)], exchange=lambda mine: [0] * world)
        arena.peers = [arena.local] * world
        args = arena.args(None)
        fused = lambda i: ops.gemm_allreduce(a, ws_[i % NB], bias, res, out, args, workspace=wsp)
        gfn, gn = graphed(fused)
        r["fused"] = timeit(gfn, n=8) / gn
    print(f"{label:18s} M={M} N={N} K={K} world={world}: " + "  ".join(f"{k} {v:7.1f} us" for k, v in r.items()), flush=True)
    for name, fn in (("plain", plain), ("fused", fused if "fused" in r else None)):
        if fn is None:
            continue
        torch.cuda.synchronize()
        for i in range(64):      # fill the 64-launch trace ring with this kernel
            fn(i)
        torch.cuda.synchronize()
        t = np.ctypeslib.as_array(cdll.lia_debug_gemm_trace(), shape=(64 * 512 * 16,)).reshape(64, 512, 16).astype(np.int64)
        for li in (10, 11):
            tt = t[li, :148]
            tt = tt[tt[:, 0] > 0]
            ent = tt[:, 0].min()
            own = tt[:, 6] > 0
            rel = lambda c, sel=own: (tt[sel, c] - ent) / 1e3
            msg = (f"   {name} launch {li}: CTAs {len(tt)} owners {own.sum()}  setup {rel(1, slice(None)).mean():5.1f}  first-full {rel(3, slice(None)).mean():5.1f}  last-MMA {rel(4, slice(None)).mean():5.1f}"
                   f"  acc-ready {rel(5, slice(None)).mean():5.1f}  pieces {rel(6).mean():5.1f}")
            if name == "fused":
                msg += f"  push-batch1 {rel(9).mean():5.1f}  pushed {rel(8).mean():5.1f}  own-rows {rel(10).mean():5.1f}  reduced {rel(11).mean():5.1f} (max {rel(11).max():5.1f})"
            msg += f"  exit mean {rel(7, slice(None)).mean():5.1f} max {rel(7, slice(None)).max():5.1f}"
            print(msg, flush=True)
    if "fused" in r:
        arena.close()
