"""One profiled launch of the fused decode exchange at the per-rank shapes of a WORLD-way OPT-30B on ONE GPU (self-loop arena,
nobody waited for: see scripts/tp_sim_trace.py), plus the plain projection of the same shape, for
  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2_prof_tp_decode_sim \
      python scripts/ncu_tp_decode_sim.py 8"""
import os, sys
os.environ["LIA_TP_SELF_LOOP"] = "1"
os.environ.setdefault("LIA_TP_NO_WAIT", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lia_b200  # noqa: F401
from lia_b200 import _lib, ops, tp

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
lib = _lib.load()
BF16 = torch.bfloat16
h, f, B = 7168, 28672, 64
for (M, N, K, label) in [(B, h, h // world, "out_proj"), (B, h, f // world, "fc2")]:
    ws_ = [(torch.randn(N, K, device=dev) * 0.02).to(BF16) for _ in range(3)]
    a = torch.randn(M, K, device=dev).to(BF16)
    bias = torch.randn(N, device=dev).to(BF16)
    res = torch.randn(M, N, device=dev).to(BF16)
    out = torch.empty(M, N, device=dev, dtype=BF16)
    wsp = ops.GemmWorkspace(ops.GemmWorkspace.bytes_for([(M, N, K)]), dev)
    arena = tp.PeerArena(0, world, dev, lib.lia_tp_recv_bytes(M, N, K, world), [("pad", 256)], exchange=lambda mine: [0] * world)
    arena.peers = [arena.local] * world
    args = arena.args(None)
    plain = lambda i: ops.gemm(a, ws_[i % 3], bias, out=out, epilogue=ops.EPI_BIAS_RESIDUAL, residual=res, workspace=wsp)
    fused = lambda i: ops.gemm_allreduce(a, ws_[i % 3], bias, res, out, args, workspace=wsp)
    for fn in (plain, fused):
        for i in range(4):
            fn(i)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        fn(2)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    print("profiled", label, flush=True)
    arena.close()
