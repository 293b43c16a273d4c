"""torchrun -n WORLD scripts/tp_decode_probe.py [layers]: decode ms/step of an OPT-30B-shaped stack (no tracing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import lia_b200
from lia_b200 import tp
from lia_b200.modeling_opt import get_config
SIM = int(os.environ.get("LIA_TP_SIM_WORLD", "0"))
if SIM:
    os.environ["LIA_TP_SELF_LOOP"] = "1"
    os.environ["LIA_TP_NO_WAIT"] = "1"
    rank, world = 0, SIM
else:
    rank, world = tp.init_from_env("nccl")
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
torch.cuda.set_device(dev)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = get_config(sys.argv[2] if len(sys.argv) > 2 else "opt-30b")
cfg.num_hidden_layers = L
m = lia_b200.OPTForCausalLM(cfg, dev, tp_rank=rank, tp_world=world).init_weights(seed=0)
B, S, new = 64, 256, 16
ids = torch.randint(3, cfg.vocab_size, (B, S), generator=torch.Generator().manual_seed(1))
kw = dict(max_new_tokens=new, min_new_tokens=new, num_minibatch=2)
best = 1e9
pre = 1e9
for i in range(6):
    m.generate(ids, **kw)
    if i >= 2:
        best = min(best, 1e3 * sum(m.last_timing["decode_s"]) / (new - 1))
        pre = min(pre, 1e3 * m.last_timing["prefill_s"])
if rank == 0:
    keys = ("LIA_TP_FUSED", "LIA_PDL", "LIA_TP_LATE_TRIGGER", "LIA_TP_POLL_BACKOFF", "LIA_TP_NO_WAIT", "LIA_TP_NO_PUSH")
    print("PROBE " + " ".join(f"{k}={os.environ.get(k, '-')}" for k in keys) + f" world={world} L={L}: decode {best:.3f} ms/step ({best / L * 1e3:.1f} us/layer)  prefill {pre:.1f} ms", flush=True)
sys.stdout.flush()
os._exit(0)
