import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, lia_b200
from lia_b200 import ops
M, N, K = 64, 7168, 7168
w = torch.randn(N, K, device="cuda").to(torch.bfloat16) * 0.02
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
bias = torch.randn(N, device="cuda").to(torch.bfloat16)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
wsp = ops.GemmWorkspace(ops.GemmWorkspace.bytes_for([(M, N, K)]), "cuda")
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n
eager = t(lambda: ops.gemm(a, w, bias, out=out, workspace=wsp))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(50): ops.gemm(a, w, bias, out=out, workspace=wsp)
graph = t(lambda: g.replay(), n=5) / 50
x = torch.randn(64, 7168, device="cuda").to(torch.bfloat16); lw = torch.ones(7168, device="cuda", dtype=torch.bfloat16); y = torch.empty_like(x)
ln_eager = t(lambda: ops.layernorm(x, lw, lw, out=y))
g2 = torch.cuda.CUDAGraph()
with torch.cuda.graph(g2):
    for _ in range(50): ops.layernorm(x, lw, lw, out=y)
ln_graph = t(lambda: g2.replay(), n=5) / 50
print(f"LIA_GEMM_DEBUG={os.environ.get('LIA_GEMM_DEBUG','0')}: gemm eager {eager:.1f} us/launch, in-graph {graph:.1f} us/launch; layernorm eager {ln_eager:.1f}, in-graph {ln_graph:.1f}")
