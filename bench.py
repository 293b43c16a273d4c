#!/usr/bin/env python
"""Headline benchmark: OPT-30B tokens/s at batch 64, 256 in / 32 out (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one generate() pass over one batch of synthetic prompts (prefill of B*S tokens +
max_new_tokens greedy tokens through every decoder layer, final LayerNorm, lm_head, argmax).
  value     tokens/s with the prompt ids already resident in HBM (CUDA events, max over ranks)
  e2e       the same metric through the public generate() call with HOST (pinned) prompt ids in and
            HOST token ids out -- the H2D/D2H copies are inside the timed region
  roofline  dominant kernel (the tcgen05 GEMM in prefill): algorithmic FLOPs / CUDA-event time of
            its launches, against the measured cuBLAS bf16 peak in MEASURED_PEAKS.json
  roofline_decode  decode step: algorithmic bytes (weights + KV + lm_head) / step time vs measured HBM GB/s
  cpu_baseline     the reference's full-CPU policy (1/1), restated (oracle/opt_ref.py), on the host cores
N > 1 runs the same workload tensor-parallel (strong scaling) with one NCCL all-reduce after each
row-parallel projection.  `--impl reference` times only the CPU restatement (rank 0).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="opt-30b")
    ap.add_argument("--batch-size", type=int, default=64)
    ap.add_argument("--input-tokens", type=int, default=256)
    ap.add_argument("--max-new-tokens", type=int, default=32)
    ap.add_argument("--num-minibatch", type=int, default=2)
    ap.add_argument("--gpu-percentage", type=int, default=100)
    ap.add_argument("--layers", type=int, default=0, help="debug: override depth (result is then NOT the headline config)")
    ap.add_argument("--weights", default="normal", choices=["normal", "dummy"],
                    help="normal(0, 0.02) init (lia/modeling_opt.py:895-904) or the reference's dummy U[0,1) weights "
                         "(utils/opt-weight-gen.py:61-62; BASELINE.json configs[4])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def committed_traffic(pair):
    """DRAM read+write bytes per launch of the prefill GEMM from the newest committed `ncu --set full` capture
    (profiles/r*_traffic.json).  Prefers the capture of the kernel variant that is running (CTA pair or one-CTA);
    says so when only the other variant has been captured."""
    import glob
    import re
    want, other = ("pair", "one-cta") if pair else ("one-cta", "pair")
    found = {}
    for tj in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json"))):
        for name, rec in json.load(open(tj)).items():
            m = re.search(r"lia_gemm_tcgen05_kernel<([^>]*)>", name)
            if not m or "dram_bytes_per_launch_avg" not in rec:
                continue
            # template arguments <SWAP, BN, STAGES[, TP, PAIR]> as ncu prints them ("0, 256, 4" / "(bool)0, (int)256, ...")
            a = [1 if x == "true" else 0 if x == "false" else int(x)
                 for x in re.findall(r"\b(true|false|\d+)\b", re.sub(r"\([a-z ]+\)", "", m.group(1)))]
            if len(a) < 3 or a[0] != 0 or a[1] != 256:
                continue                                    # decode (swap-AB) or narrow-N variants
            key = "pair" if (len(a) >= 5 and a[4] == 1) else "one-cta"
            found[key] = (rec["dram_bytes_per_launch_avg"], os.path.basename(tj), name)
    if want in found:
        v, f, name = found[want]
        return v, f"ncu --set full capture of {name} ({f})"
    if other in found:
        v, f, name = found[other]
        return v, (f"from the ncu capture of {name} ({f}): the other prefill variant "
                   f"({'one-CTA cta_group::1' if pair else 'CTA-pair cta_group::2'} kernel); the running variant has no committed capture yet")
    return None, None


class ClockSampler:
    """SM clock + throttle reasons during the timed region (NVML; nvidia-smi as a fallback)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_baseline(args, cfg, steps=1, warmup=0):
    """Reference full-CPU policy (prefill-policy 1 / decoding-policy 1), restated with stock PyTorch
    CPU ops (oracle/opt_ref.py: CpuPolicy1Runner) on a bounded sample of the workload."""
    import torch
    from oracle.opt_ref import CpuPolicy1Runner
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    h, H, f, L = cfg.hidden_size, cfg.num_attention_heads, cfg.ffn_dim, cfg.num_hidden_layers
    B, S, new = args.batch_size, args.input_tokens, args.max_new_tokens
    # probe the host's bf16 GEMM rate to size the sample to roughly 10-30 s
    a = torch.randn(512, h).to(torch.bfloat16)
    w = torch.randn(h, h).to(torch.bfloat16)
    torch.nn.functional.linear(a, w)
    t0 = time.perf_counter()
    torch.nn.functional.linear(a, w)
    tf = 2 * 512 * h * h / (time.perf_counter() - t0) / 1e12
    layer_prefill_flops = 24.0 * h * h * B * S
    Bs = B
    while Bs > 1 and layer_prefill_flops * (Bs / B) / (tf * 1e12) > 12.0:
        Bs //= 2
    r = CpuPolicy1Runner(h, H, f, 1, Bs, S + new)
    times = []
    for i in range(warmup + steps):
        t = r.run(S, new)
        if i >= warmup:
            times.append(t)
    t_layer = sum(times) / len(times)
    # the steps either side of the stack, once per generated token: final LayerNorm, tied lm_head on the last position,
    # argmax (models.py:423-431, greedy_search.py:395) -- timed on their own and added `new` times
    V = cfg.vocab_size
    e = (torch.randn(V, h) * 0.02).to(torch.bfloat16)
    xh = torch.randn(Bs, h).to(torch.bfloat16)
    lw, lb = torch.ones(h, dtype=torch.bfloat16), torch.zeros(h, dtype=torch.bfloat16)
    with torch.inference_mode():
        th = []
        for i in range(3):
            t0 = time.perf_counter()
            torch.argmax(torch.nn.functional.linear(torch.nn.functional.layer_norm(xh, (h,), lw, lb, 1e-5), e).float(), dim=-1)
            th.append(time.perf_counter() - t0)
    t_head = min(th[1:])
    t_total = t_layer * L + t_head * new
    tok_s = (Bs * new) / t_total
    sample = (f"1 of {L} decoder layers of {cfg.name} at batch {Bs} (of {B}), input {S}, {new} new tokens, layer time "
              f"extrapolated x{L} layers, plus {new} x (final LayerNorm + lm_head + argmax) timed separately "
              f"({t_head * 1e3:.1f} ms each); embedding lookups excluded; host GEMM probe {tf:.2f} TFLOP/s")
    return {"value": tok_s, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample,
            "sample_seconds": t_layer, "head_seconds": t_head}, t_total / L


def main():
    args = parse()
    # stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version banner) go to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    try:
        return _main(args, real_stdout)
    finally:
        real_stdout.flush()


def _main(args, json_out):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import lia_b200
    from lia_b200.modeling_opt import get_config
    cfg = get_config(args.model)
    if args.layers:
        cfg.num_hidden_layers = args.layers
    B, S, new = args.batch_size, args.input_tokens, args.max_new_tokens
    workload = (f"{cfg.name} bf16 {'random-init' if args.weights == 'normal' else 'dummy U[0,1) weights'}, {'fully HBM-resident' if args.gpu_percentage >= 100 else f'gpu-percentage {args.gpu_percentage}, rest streamed from pinned host'}, "
                f"batch {B}, input {S}, max-new-tokens {new}, num-minibatch {args.num_minibatch}"
                + (f" [DEBUG depth {args.layers}: not the headline config]" if args.layers else
                   " (BASELINE.json configs[1])" if (args.model, B, S, new, args.gpu_percentage) == ("opt-30b", 64, 256, 32, 100) else
                   " (BASELINE.json configs[3])" if (args.model, B, S, new) == ("opt-66b", 64, 512, 64) else
                   " (BASELINE.json configs[4])" if (args.model, B, S, new, args.weights) == ("opt-175b", 64, 256, 32, "dummy") else ""))
    config = {"workload": workload, "parallelism": f"tp{world}", "l2": "inputs_exceed_l2 (weights+KV per step >> 126 MB)",
              "cuda_graphs": not args.no_graphs}

    if args.impl == "reference":
        if rank != 0:
            return 0
        cb, t = cpu_baseline(args, cfg, steps=max(1, args.steps), warmup=min(args.warmup, 1))
        line = {"impl": "reference", "metric": "tokens/s", "value": cb["value"], "unit": "tokens/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * cfg.num_hidden_layers * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": config, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), file=json_out)
        return 0

    import torch
    from lia_b200 import _lib, ops, tp
    rank, world = tp.init_from_env("nccl")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pk = peaks()

    m = lia_b200.OPTForCausalLM(cfg, dev, tp_rank=rank, tp_world=world)
    m.use_cuda_graphs = not args.no_graphs
    m.init_weights(seed=0, kind=args.weights, gpu_percentage=args.gpu_percentage)
    g = torch.Generator().manual_seed(1234)
    ids_host = torch.randint(3, cfg.vocab_size, (B, S), generator=g).pin_memory()
    ids_dev = ids_host.to(dev)
    kw = dict(max_new_tokens=new, min_new_tokens=new, do_sample=False, num_beams=1, prefill_policy=0, decoding_policy=0,
              gpu_percentage=args.gpu_percentage, num_minibatch=args.num_minibatch, pin_weight=True)

    # warm-up (>= 3: eager, graph capture, graph replay); count our kernel launches on the eager pass
    c0 = _lib.launch_count
    m.generate(ids_dev, **kw)
    launches_per_step = _lib.launch_count - c0
    for _ in range(max(args.warmup, 3) - 1):
        m.generate(ids_dev, **kw)

    def timed(inp):
        tp.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        prefill, decode = [], []
        e0.record()
        for _ in range(args.steps):
            out = m.generate(inp, **kw)
            prefill.append(m.last_timing["prefill_s"])
            decode += m.last_timing["decode_s"]
        e1.record()
        tp.barrier()
        torch.cuda.synchronize(dev)
        sec = tp.max_over_ranks(e0.elapsed_time(e1) / 1e3, dev)
        return sec, out, prefill, decode

    with ClockSampler(local) as clk:
        sec, out, prefill, decode = timed(ids_dev)
    sec_e2e, out_h, _, _ = timed(ids_host)
    assert out_h.device.type == "cpu" and out_h.shape == (B, S + new)
    tok = B * new * args.steps
    value, e2e = tok / sec, tok / sec_e2e
    t_prefill = sum(prefill) / len(prefill)
    t_decode = sum(decode) / max(1, len(decode))

    # ---- roofline of the dominant kernel: instrument one prefill pass, CUDA events around every GEMM launch
    L, h, f, V = cfg.num_hidden_layers, cfg.hidden_size, cfg.ffn_dim, cfg.vocab_size
    recs = []
    orig = ops.gemm

    def timed_gemm(a, w, *a2, **k2):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        r = orig(a, w, *a2, **k2)
        e.record()
        recs.append((a.shape[0], w.shape[0], a.shape[1], s, e))
        return r

    ops.gemm = timed_gemm
    st = next(iter(m._states.values()))
    m._prefill(st, args.num_minibatch, -1)
    ops.gemm = orig
    torch.cuda.synchronize(dev)
    big = [(M_, N_, K_, s.elapsed_time(e)) for (M_, N_, K_, s, e) in recs if M_ > 128]
    roof = None
    if big:
        flops = sum(2.0 * M_ * N_ * K_ for M_, N_, K_, _ in big)
        ms = sum(t for *_, t in big)
        ach = flops / (ms / 1e3) / 1e12
        traffic, traffic_note = None, None
        if not args.layers and world == 1:
            traffic, traffic_note = committed_traffic(os.environ.get("LIA_GEMM_2CTA", "1") != "0")
        roof = {"kernel": "lia_gemm_tcgen05_kernel (prefill projections)", "bound": "tensor", "achieved": ach,
                "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": ach / pk["bf16_sustained"], "traffic": traffic,
                "frac_of_burst_peak": ach / pk["bf16_burst"],
                "launches": len(big), "avg_launch_ms": ms / len(big), "flops_per_launch": flops / len(big),
                "peak_source": pk["source"] + ", sustained figure (kernel timed inside a long step); burst "
                + f"{pk['bf16_burst']}", "share_of_step": (ms / 1e3) / (sec / args.steps),
                "traffic_note": traffic_note}
    # decode step: algorithmic bytes = L*W_l + L*4*B*T*h + 2*h*V   (SURVEY.md 8d), per rank
    Wl = 2.0 * (12 * h * h + 13 * h)
    Tavg = S + (new - 1) / 2.0 + 0.5
    dec_bytes = (L * Wl + L * 4.0 * B * Tavg * h) / world + 2.0 * h * V
    roof_dec = {"kernel": "decode step (swap-AB tcgen05 GEMMs + flash-decoding attention + LN, CUDA graph)", "bound": "hbm",
                "achieved": dec_bytes / t_decode / 1e9 if t_decode > 0 else None, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": (dec_bytes / t_decode / 1e9) / pk["hbm_gbs"] if t_decode > 0 else None, "traffic": None,
                "bytes_per_step": dec_bytes, "ms_per_decode_step": t_decode * 1e3, "peak_source": pk["source"]}

    if rank == 0:
        line = {"metric": "tokens/s", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
                "clocks": clk.summary(),
                "e2e": {"value": e2e, "unit": "tokens/s", "h2d_bytes_per_step": B * S * 8, "d2h_bytes_per_step": B * (S + new) * 8},
                "gpu_launches": launches_per_step * args.steps,
                "prefill_ms": t_prefill * 1e3, "decode_ms_per_step": t_decode * 1e3,
                "roofline": roof, "roofline_decode": roof_dec}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"], _ = cpu_baseline(args, cfg)
        print(json.dumps(line), file=json_out)
        json_out.flush()
    tp.barrier()
    return 0


if __name__ == "__main__":
    sys.exit(main())
