"""One profiled launch of the prefill attention at the headline shape (OPT-30B, one minibatch: B=32, H=56, S=256, d=128) and at
the reference's long-prompt shape (B=4, S=2016) for `ncu --profile-from-start off`.
  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2_prof_attn_prefill python scripts/ncu_attn_prefill.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lia_b200  # noqa: F401
from lia_b200 import ops

dev, BF16 = "cuda", torch.bfloat16
for (B, H, S, d) in [(32, 56, 256, 128), (4, 56, 2016, 128)]:
    q = (torch.randn(B * S, H * d, device=dev) * 0.3).to(BF16)
    kc = torch.randn(S, B, H, d, device=dev).to(BF16)
    vc = torch.randn(S, B, H, d, device=dev).to(BF16)
    out = torch.empty(B * S, H * d, device=dev, dtype=BF16)
    for _ in range(2):
        ops.attn_prefill(q, kc, vc, B, S, 0, out=out)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    ops.attn_prefill(q, kc, vc, B, S, 0, out=out)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled", B, H, S, d, flush=True)
