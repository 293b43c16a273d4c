"""torchrun -n WORLD scripts/tp_timeline.py [layers]: GEMM-level timeline of one tensor-parallel decode step
(LIA_GEMM_TRACE stamps inside CUDA-graph replay): span of every projection and the gap before it
(LayerNorm / attention / all-reduce kernels live in the gaps).  Run with LIA_TP_FUSED=1 and =0."""
import ctypes, os, sys
os.environ["LIA_GEMM_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import lia_b200
from lia_b200 import _lib, tp
from lia_b200.modeling_opt import get_config

SIM = int(os.environ.get("LIA_TP_SIM_WORLD", "0"))    # one process pretending to be rank 0 of SIM (timing probe, no real exchange)
if SIM:
    os.environ["LIA_TP_SELF_LOOP"] = "1"
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
L = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = get_config("opt-30b")
cfg.num_hidden_layers = L
m = lia_b200.OPTForCausalLM(cfg, dev, tp_rank=rank, tp_world=world).init_weights(seed=0)
B, S, new = 64, 256, 4
ids = torch.randint(3, cfg.vocab_size, (B, S), generator=torch.Generator().manual_seed(1))
kw = dict(max_new_tokens=new, min_new_tokens=new, num_minibatch=2)
for _ in range(4):
    m.generate(ids, **kw)
torch.cuda.synchronize(); dist.barrier()
cdll = ctypes.CDLL(_lib.LIB_PATH)
cdll.lia_debug_gemm_trace.restype = ctypes.POINTER(ctypes.c_ulonglong)
t = np.ctypeslib.as_array(cdll.lia_debug_gemm_trace(), shape=(64 * 512 * 16,)).reshape(64, 512, 16).astype(np.int64)
rows = []
for s in range(63):
    tt = t[s, :148]
    live = tt[:, 0] > 0
    if live.any():
        own = tt[:, 8] > 0
        det = None
        if own.any():      # fused projection: owner-CTA phases relative to the launch's first CTA entry
            e = tt[live, 0].min()
            det = [float((tt[own, c] - e).mean()) / 1e3 for c in (5, 6, 8, 11, 7)] + [float((tt[own, 11] - e).max()) / 1e3]
        rows.append((tt[live, 0].min(), tt[live, 7].max(), int(live.sum()), det))
rows.sort()
rows = rows[-(4 * L + 1):]          # the last decode step: 4 projections per layer + lm_head
names = ["qkv", "out", "fc1", "fc2"]
names = ["qkv", "out", "fc1", "fc2"]
for rr in range(1 if SIM else world):
    dist.barrier()
    if rr == rank:
        print(f"rank {rank} PDL={os.environ.get('LIA_PDL', '1')} fused={os.environ.get('LIA_TP_FUSED', '1')} late={os.environ.get('LIA_TP_LATE_TRIGGER', '-')} "
              f"world={world} decode ms/step {1e3 * sum(m.last_timing['decode_s']) / (new - 1):.3f}")
        t0 = rows[0][0]
        per = {n: [0.0, 0.0] for n in names}
        for i, (a, b, n, det) in enumerate(rows[:-1]):
            gap = (a - rows[i - 1][1]) / 1e3 if i else 0.0
            nm = names[i % 4]
            per[nm][0] += (b - a) / 1e3
            per[nm][1] += gap
            if 4 <= i < 12:
                extra = ""
                if det is not None and nm in ("out", "fc2") and os.environ.get("LIA_TP_FUSED", "1") != "0":
                    extra = f"  owners: acc {det[0]:5.1f} pieces {det[1]:5.1f} pushed {det[2]:5.1f} reduced {det[3]:5.1f} (max {det[5]:5.1f})"
                print(f"  {nm:4s} entry {(a - t0) / 1e3:7.1f}  exit {(b - t0) / 1e3:7.1f}  span {(b - a) / 1e3:6.1f}  gap before {gap:6.1f}{extra}")
        for nm in names:
            print(f"  avg {nm:4s}: span {per[nm][0] / L:6.1f} us   gap before {per[nm][1] / L:6.1f} us")
        print(f"  step (first GEMM entry -> lm_head exit): {(rows[-1][1] - rows[0][0]) / 1e3:8.1f} us", flush=True)
dist.barrier()
sys.stdout.flush()
os._exit(0)
