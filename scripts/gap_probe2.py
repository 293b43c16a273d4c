import os, sys, ctypes
os.environ["LIA_GEMM_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np, lia_b200
from lia_b200 import ops, _lib
_lib.load()
cdll = ctypes.CDLL(_lib.LIB_PATH)
cdll.lia_debug_gemm_trace.restype = ctypes.POINTER(ctypes.c_ulonglong)
cdll.lia_debug_marker.argtypes = [ctypes.c_int, ctypes.c_void_p]
M, N, K = 64, 7168, 7168
ws_ = [torch.randn(N, K, device="cuda").to(torch.bfloat16) * 0.02 for _ in range(3)]
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
bias = torch.randn(N, device="cuda").to(torch.bfloat16)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
wsp = ops.GemmWorkspace(ops.GemmWorkspace.bytes_for([(M, N, K)]), "cuda")
st = torch.cuda.current_stream().cuda_stream
ops.gemm(a, ws_[0], bias, out=out, workspace=wsp); torch.cuda.synchronize()   # seq 0
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    s2 = torch.cuda.current_stream().cuda_stream
    for i in range(6):
        cdll.lia_debug_marker(2 * i, s2)
        ops.gemm(a, ws_[i % 3], bias, out=out, workspace=wsp)    # seq 1..6
        cdll.lia_debug_marker(2 * i + 1, s2)
for _ in range(3): g.replay()
torch.cuda.synchronize()
t = np.ctypeslib.as_array(cdll.lia_debug_gemm_trace(), shape=(64 * 1024 * 8,)).astype(np.int64).reshape(64, 1024, 8)
mk = t[63].reshape(-1)
for i in range(6):
    tt = t[1 + i, :148]
    m0, m1 = mk[2 * i], mk[2 * i + 1]
    ent, ext = tt[:, 0].min(), tt[:, 7].max()
    print(f"gemm {i}: marker_before -> first CTA entry {(ent-m0)/1e3:6.2f} us | CTA span {(ext-ent)/1e3:6.2f} us | last CTA exit -> marker_after {(m1-ext)/1e3:6.2f} us | marker to marker {(m1-m0)/1e3:6.2f} us")
