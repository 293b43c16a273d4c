"""Per-CTA timeline of the decode GEMM (LIA_GEMM_TRACE=1): where do the microseconds go?"""
import os, sys, ctypes
os.environ["LIA_GEMM_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import lia_b200
from lia_b200 import ops, _lib
lib = _lib.load()
cdll = ctypes.CDLL(_lib.LIB_PATH)
cdll.lia_debug_gemm_trace.restype = ctypes.POINTER(ctypes.c_ulonglong)
names = ["entry", "setup done", "first TMA issued", "first stage full", "last MMA committed", "epi: last acc ready", "epi: pieces ready", "exit"]
def run(M, N, K, epi, label, n=12):
    ws_ = [torch.randn(N, K, device="cuda").to(torch.bfloat16) * 0.02 for _ in range(3)]
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    bias = torch.randn(N, device="cuda").to(torch.bfloat16)
    res = torch.randn(M, N, device="cuda").to(torch.bfloat16) if epi == 2 else None
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    wsp = ops.GemmWorkspace(ops.GemmWorkspace.bytes_for([(M, N, K)]), "cuda")
    global seq
    torch.cuda.synchronize()
    first = seq
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        ops.gemm(a, ws_[i % 3], bias, out=out, epilogue=epi, residual=res, workspace=wsp)
    e1.record()
    seq += n
    torch.cuda.synchronize()
    ptr = cdll.lia_debug_gemm_trace()
    t = np.ctypeslib.as_array(ptr, shape=(64 * 512 * 16,)).reshape(64, 512, 16).astype(np.int64)
    print(f"--- {label} M={M} N={N} K={K}: event time {e0.elapsed_time(e1)*1e3/n:.1f} us/launch")
    prev_exit = None
    for i in range(first, first + n):
        tt = t[i % 64, :148]
        ent, ext = tt[:, 0].min(), tt[:, 7].max()
        ff = tt[:, 3]
        gap = (ent - prev_exit) / 1e3 if prev_exit is not None else float("nan")
        print(f"   launch {i-first:2d}: span {(ext-ent)/1e3:6.1f} us   gap before {gap:6.1f} us   entry spread {(tt[:,0].max()-ent)/1e3:5.2f}  first-full avg {(ff.mean()-ent)/1e3:5.2f}  exit spread {(ext-tt[:,7].min())/1e3:5.2f}")
        prev_exit = ext
seq = 0
h, f = 7168, 28672
run(64, 3 * h, h, 0, "qkv")
run(64, h, h, 2, "out")
run(64, f, h, 1, "fc1")
run(64, h, f, 2, "fc2")
