"""torchrun -n 2 scripts/tp_chain_probe.py: per-iteration time of graph-captured chains [projection, LayerNorm] at the
decode out_proj shape -- does a kernel that stored into peer memory delay the START of its successor?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import lia_b200
from lia_b200 import _lib, ops, tp
SIM = int(os.environ.get("LIA_TP_SIM_WORLD", "0"))
if SIM:
    os.environ["LIA_TP_NO_WAIT"] = "1"
    rank, world = 0, SIM
    class _D:
        @staticmethod
        def barrier(): pass
    dist = _D()
else:
    rank, world = tp.init_from_env("nccl")
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
torch.cuda.set_device(dev)
lib = _lib.load()
BF16 = torch.bfloat16
h, B = 7168, 64
M, N, K = B, h, h // world
NB = 4
ws_ = [(torch.randn(N, K, device=dev) * 0.02).to(BF16) for _ in range(NB)]
a = torch.randn(M, K, device=dev).to(BF16)
bias = torch.randn(N, device=dev).to(BF16)
res = torch.randn(M, N, device=dev).to(BF16)
lnw, lnb = torch.ones(N, device=dev, dtype=BF16), torch.zeros(N, device=dev, dtype=BF16)
arena = tp.PeerArena(rank, world, dev, lib.lia_tp_recv_bytes(M, N, K, world), [("out", M * N * 2)],
                     exchange=(lambda mine: [0] * world) if SIM else None)
if SIM:
    arena.peers = [arena.local] * world
w2 = [(torch.randn(3 * N // world, N, device=dev) * 0.02).to(BF16) for _ in range(NB)]     # the qkv projection that follows
b2 = torch.randn(3 * N // world, device=dev).to(BF16)
o2 = torch.empty(M, 3 * N // world, device=dev, dtype=BF16)
wsp2 = ops.GemmWorkspace(ops.GemmWorkspace.bytes_for([(M, 3 * N // world, N)]), dev)
nxt = lambda i: ops.gemm(lno, w2[i % NB], b2, out=o2, epilogue=ops.EPI_BIAS, workspace=wsp2)
nxt_shared = lambda i: ops.gemm(lno, w2[i % NB], b2, out=o2, epilogue=ops.EPI_BIAS, workspace=wsp)
out = torch.empty(M, N, device=dev, dtype=BF16)
lno = torch.empty(M, N, device=dev, dtype=BF16)
wsp = ops.GemmWorkspace(ops.GemmWorkspace.bytes_for([(M, N, K)]), dev)
args = arena.args(None)
plain = lambda i: ops.gemm(a, ws_[i % NB], bias, out=out, epilogue=ops.EPI_BIAS_RESIDUAL, residual=res, workspace=wsp)
fused = lambda i: ops.gemm_allreduce(a, ws_[i % NB], bias, res, out, args, workspace=wsp)
ln = lambda i: ops.layernorm(out, lnw, lnb, out=lno)

def chain(fns, n=16):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            for f in fns:
                f(i)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize(); dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        g.replay()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / 10 / n * 1e3

r = {"plain": chain([plain]), "plain+ln": chain([plain, ln]), "fused": chain([fused]), "fused+ln": chain([fused, ln]), "ln": chain([ln]),
     "plain+ln+qkv": chain([plain, ln, nxt]), "fused+ln+qkv": chain([fused, ln, nxt]), "qkv": chain([nxt])}
if rank == 0:
    keys = ("LIA_TP_LATE_TRIGGER", "LIA_TP_NO_WAIT", "LIA_TP_NO_PUSH", "LIA_TP_OPTS")
    print("CHAIN " + " ".join(f"{k}={os.environ.get(k, '-')}" for k in keys) + ": " + "  ".join(f"{k} {v:6.1f}" for k, v in r.items()), flush=True)
sys.stdout.flush()
os._exit(0)
