"""Per-operation timeline of one decode-program launch (LIA_PROGRAM_DEBUG=1): for CTA 0, G/2 and G-1, %globaltimer when the
producer / MMA / compute warps reach each operation and when the compute warps finish it.  Prints per-kind averages:
  wait  = compute warps: op start (after dependency) minus previous op done     busy = op done minus op start
  python scripts/program_timeline.py [layers=4]"""
import ctypes, os, sys
os.environ["LIA_PROGRAM_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lia_b200
from lia_b200.modeling_opt import get_config

L = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cfg = get_config("opt-30b"); cfg.num_hidden_layers = L
lib = lia_b200._lib.load()
f = lib.lia_debug_program_progress; f.restype = ctypes.POINTER(ctypes.c_int)
base = f()
m = lia_b200.OPTForCausalLM(cfg, "cuda").init_weights(seed=0)
ids = torch.randint(3, cfg.vocab_size, (64, 256), generator=torch.Generator().manual_seed(1))
for _ in range(2):
    m.generate(ids, max_new_tokens=8, min_new_tokens=8)
torch.cuda.synchronize()
st = next(iter(m._states.values()))
n_ops = st.program.num_ops
tl = ctypes.cast(ctypes.addressof(base.contents) + 4096 * 4, ctypes.POINTER(ctypes.c_uint64))
kinds = ["embed"] + ["ln1", "qkv", "attn", "out", "ln2", "fc1", "fc2"] * L + ["lnf", "lm_head", "argmax"]
assert len(kinds) == n_ops, (len(kinds), n_ops)
for slot, name in enumerate(["cta 0", "cta G/2", "cta G-1"]):
    T = [[tl[((slot * 4096) + i) * 4 + r] for r in range(4)] for i in range(n_ops)]
    t0 = T[0][2]
    agg = {}
    for i in range(1, n_ops):
        wait = (T[i][2] - T[i - 1][3]) / 1e3
        busy = (T[i][3] - T[i][2]) / 1e3
        lead = (T[i][2] - T[i][0]) / 1e3       # how far ahead of the compute warps the producer reached this op
        a = agg.setdefault(kinds[i], [0, 0.0, 0.0, 0.0]); a[0] += 1; a[1] += wait; a[2] += busy; a[3] += lead
    total = (T[n_ops - 1][3] - t0) / 1e3
    print(f"{name}: whole program {total:.1f} us over {n_ops} ops")
    for k, (n, w, b, l) in agg.items():
        print(f"   {k:8s} x{n:3d}: wait {w / n:7.2f} us   busy {b / n:7.2f} us   producer lead {l / n:7.2f} us")
