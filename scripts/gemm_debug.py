"""GPU debugging aid: run the tcgen05 GEMM on a few shapes and print error patterns."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lia_b200
from lia_b200 import ops

torch.manual_seed(0)
def run(M, N, K, epi=0):
    a = (torch.randn(M, K)).to(torch.bfloat16).cuda()
    w = (torch.randn(N, K) * K ** -0.5).to(torch.bfloat16).cuda()
    b = torch.randn(N).to(torch.bfloat16).cuda()
    ws = ops.GemmWorkspace(ops.GemmWorkspace.bytes_for([(M, N, K)]), "cuda")
    y = ops.gemm(a, w, b, epilogue=epi, workspace=ws)
    torch.cuda.synchronize()
    ref = (a.float() @ w.float().t()).to(torch.bfloat16).float() + b.float()
    ref = ref.to(torch.bfloat16).float()
    if epi == 1: ref = torch.relu(ref)
    err = (y.float() - ref).abs()
    bad = err > 0.05 * ref.abs().max()
    print(f"M={M} N={N} K={K} epi={epi}: max err {err.max().item():.4g} ref max {ref.abs().max().item():.3g} bad {int(bad.sum())}/{bad.numel()}", flush=True)
    if bad.any():
        rows = torch.nonzero(bad.any(1)).flatten().tolist()
        cols = torch.nonzero(bad.any(0)).flatten().tolist()
        print("   bad rows (first 20):", rows[:20], "count", len(rows))
        print("   bad cols (first 20):", cols[:20], "count", len(cols))
        print("   y[0,:8]", y[0, :8].float().tolist())
        print("   r[0,:8]", ref[0, :8].tolist())

for shp in [(256, 128, 64), (256, 256, 128), (128 * 3, 256, 64 * 5), (64, 128, 64), (64, 128, 256), (16, 256, 512), (64, 7168, 7168), (2048, 7168, 7168)]:
    try:
        run(*shp)
    except Exception as e:
        print("EXC", shp, e, flush=True)
