"""torchrun -n WORLD scripts/tp_microbench.py: decode-shape row-parallel projection, three ways
  plain   lia_gemm_bf16(BIAS_RESIDUAL) at the sharded shape (no exchange: lower bound)
  nccl    lia_gemm_bf16(BIAS) -> NCCL all_reduce -> lia_residual_add_bf16
  fused   lia_gemm_allreduce_bf16 (one kernel, partial tiles pushed over NVLink peer memory)
each timed back-to-back eagerly and as a CUDA graph of 32 calls; with LIA_GEMM_TRACE=1 also prints the
owner-CTA timeline of the fused kernel (push / fence / peer wait / reduce)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import lia_b200
from lia_b200 import _lib, ops, tp

rank, world = tp.init_from_env("nccl")
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
torch.cuda.set_device(dev)
lib = _lib.load()
BF16 = torch.bfloat16
h, f, B = 7168, 28672, 64
NB = 4


def timeit(fn, n=64, warm=8):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize(); dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(n):
        fn(i)
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


def graphed(fn, n=32):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            fn(i)
    return lambda _i: g.replay(), n


for (M, N, K, label) in [(B, h, h // world, "out_proj"), (B, h, f // world, "fc2"), (8192, h, h // world, "out_proj prefill"), (8192, h, f // world, "fc2 prefill")]:
    ws_ = [(torch.randn(N, K, device=dev) * 0.02).to(BF16) for _ in range(NB)]
    a = torch.randn(M, K, device=dev).to(BF16)
    bias = torch.randn(N, device=dev).to(BF16)
    res = torch.randn(M, N, device=dev).to(BF16)
    arena = tp.PeerArena(rank, world, dev, lib.lia_tp_recv_bytes(M, N, K, world), [("out", M * N * 2)])
    out = arena.tensor("out", (M, N))
    part = torch.empty(M, N, device=dev, dtype=BF16)
    wsp = ops.GemmWorkspace(ops.GemmWorkspace.bytes_for([(M, N, K)]), dev)
    args = arena.args(out if M > 128 else None)
    plain = lambda i: ops.gemm(a, ws_[i % NB], bias, out=out, epilogue=ops.EPI_BIAS_RESIDUAL, residual=res, workspace=wsp)
    def nccl(i):
        ops.gemm(a, ws_[i % NB], bias, out=part, epilogue=ops.EPI_BIAS, workspace=wsp)
        dist.all_reduce(part)
        ops.residual_add(part, res, out=out)
    fused = lambda i: ops.gemm_allreduce(a, ws_[i % NB], bias, res, out, args, workspace=wsp)
    n = 64 if M <= 128 else 8
    r = {}
    for name, fn in (("plain", plain), ("nccl", nccl), ("fused", fused)):
        r[name] = timeit(fn, n=n)
        gfn, gn = graphed(fn, n=32 if M <= 128 else 4)
        r[name + "_graph"] = timeit(gfn, n=8, warm=2) / gn
    arena.check()
    if rank == 0:
        print(f"{label:18s} M={M} N={N} K={K} world={world}: " + "  ".join(f"{k} {v:7.1f} us" for k, v in r.items()), flush=True)
    if os.environ.get("LIA_GEMM_TRACE") and M <= 128:
        cdll = ctypes.CDLL(_lib.LIB_PATH)
        cdll.lia_debug_gemm_trace.restype = ctypes.POINTER(ctypes.c_ulonglong)
        torch.cuda.synchronize(); dist.barrier()
        for i in range(70):      # wrap the 64-launch ring so the last launches are all fused ones
            fused(i)
        torch.cuda.synchronize()
        t = np.ctypeslib.as_array(cdll.lia_debug_gemm_trace(), shape=(64 * 512 * 16,)).reshape(64, 512, 16).astype(np.int64)
        if rank == 0:
            for li in range(3):
                tt = t[li, :148]
                own = tt[:, 8] > 0
                ent = tt[:, 0].min()
                rel = lambda c: (tt[own, c] - ent) / 1e3
                print(f"   trace launch {li}: span {(tt[:,7].max()-ent)/1e3:6.1f} us; owners {own.sum()}: "
                      f"acc-ready {rel(5).mean():6.1f}  pieces {rel(6).mean():6.1f}  pushed {rel(8).mean():6.1f}  fenced {rel(9).mean():6.1f} "
                      f"peer-ready {rel(10).mean():6.1f} (max {rel(10).max():6.1f})  reduced {rel(11).mean():6.1f}  exit max {(tt[:,7].max()-ent)/1e3:6.1f}", flush=True)
    arena.close()
dist.barrier()
dist.destroy_process_group()
