#!/usr/bin/env python
"""Headline benchmark: OPT-30B tokens/s at batch 64, 256 in / 32 out (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c1|c2|c3|c4|c5a|c5b]

A "step" is one generate() pass over one batch of synthetic prompts (prefill of B*S tokens +
max_new_tokens greedy tokens through every decoder layer, final LayerNorm, lm_head, argmax).
  value     tokens/s with the prompt ids already resident in HBM (CUDA events, max over ranks)
  e2e       the same metric through the public generate() call with HOST (pinned) prompt ids in and
            HOST token ids out -- the H2D/D2H copies are inside the timed region
  roofline  dominant kernel (the tcgen05 GEMM in prefill): algorithmic FLOPs / CUDA-event time of
            its launches, against the measured cuBLAS bf16 peak in MEASURED_PEAKS.json
  roofline_decode  decode step: algorithmic bytes (weights + KV + lm_head) / step time vs measured HBM GB/s
  parity    OUTSIDE the timed region, every N: the timed model's own outputs (last-layer hidden state after the
            prefill, the 32 greedy tokens of the last timed generate()) against the oracle run at full depth on the
            unsharded weights (rank 0) -- so a throughput line is always the throughput of a checked result
  cpu_baseline     the reference's full-CPU policy (1/1), restated (oracle/opt_ref.py), on the host cores
N > 1 runs the same workload tensor-parallel (strong scaling): QKV / fc1 column-split, out_proj / fc2 row-split with the
all-reduce fused into the projection kernel over NVLink peer memory.  `--impl reference` times only the CPU
restatement (rank 0).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json configs, by position (c2 is the headline; c5 has a 1-GPU streamed and an 8-GPU resident form)
CONFIGS = {
    "c1": dict(model="opt-1.3b", batch_size=8, input_tokens=256, max_new_tokens=32, num_minibatch=1, gpu_percentage=100, weights="normal"),
    "c2": dict(model="opt-30b", batch_size=64, input_tokens=256, max_new_tokens=32, num_minibatch=2, gpu_percentage=100, weights="normal"),
    "c3": dict(model="opt-30b", batch_size=512, input_tokens=256, max_new_tokens=32, num_minibatch=4, gpu_percentage=10, weights="normal"),
    "c4": dict(model="opt-66b", batch_size=64, input_tokens=512, max_new_tokens=64, num_minibatch=2, gpu_percentage=100, weights="normal"),
    "c5a": dict(model="opt-175b", batch_size=64, input_tokens=256, max_new_tokens=32, num_minibatch=2, gpu_percentage=20, weights="dummy"),
    "c5b": dict(model="opt-175b", batch_size=64, input_tokens=256, max_new_tokens=32, num_minibatch=2, gpu_percentage=100, weights="dummy"),
}
CONFIG_LABEL = {"c1": "BASELINE.json configs[0]", "c2": "BASELINE.json configs[1]", "c3": "BASELINE.json configs[2]",
                "c4": "BASELINE.json configs[3]", "c5a": "BASELINE.json configs[4], 1 GPU streamed", "c5b": "BASELINE.json configs[4], TP resident"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS), help="a BASELINE.json config by name (default c2 = the headline)")
    ap.add_argument("--model", default=None)
    ap.add_argument("--batch-size", type=int, default=None)
    ap.add_argument("--input-tokens", type=int, default=None)
    ap.add_argument("--max-new-tokens", type=int, default=None)
    ap.add_argument("--num-minibatch", type=int, default=None)
    ap.add_argument("--gpu-percentage", type=int, default=None)
    ap.add_argument("--layers", type=int, default=0, help="debug: override depth (result is then NOT the headline config)")
    ap.add_argument("--weights", default=None, choices=["normal", "dummy"],
                    help="normal(0, 0.02) init (lia/modeling_opt.py:895-904) or the reference's dummy U[0,1) weights "
                         "(utils/opt-weight-gen.py:61-62; BASELINE.json configs[4])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--quick", action="store_true",
                    help="PCIe-bound configs whose single generate() takes minutes (c3, c5a): ONE warm-up pass instead of >= 3 and no "
                         "second (host-buffer) timed pass; the line says so -- never use it for the headline config")
    a = ap.parse_args()
    preset = CONFIGS[a.config or "c2"]
    explicit = {k: getattr(a, k) for k in preset if getattr(a, k) is not None}
    for k, v in preset.items():
        if getattr(a, k) is None:
            setattr(a, k, v)
    a.config_name = (a.config or "c2") if all(preset[k] == v for k, v in explicit.items()) else None
    return a


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
    for tj in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")) + glob.glob(os.path.join(ROOT, "profiles", "r*", "*traffic.json"))):
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
            found[key] = (rec["dram_bytes_per_launch_avg"], os.path.relpath(tj, ROOT), name)
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


def cpu_baseline(args, cfg, steps=1, warmup=0, budget_s=20.0):
    """Reference full-CPU policy (prefill-policy 1 / decoding-policy 1), restated with stock PyTorch
    CPU ops (oracle/opt_ref.py: CpuPolicy1Runner) on a BOUNDED sample of the workload.

    The sample is the WHOLE model when one pass fits the per-step budget (opt-1.3b, BASELINE.json configs[0]: nothing
    is extrapolated); otherwise TWO consecutive decoder layers (so that the second layer's weights really come from
    DRAM, not from the cache the first one warmed) at a batch cut down until one pass fits the budget, scaled to the
    full depth, plus the per-token head (final LayerNorm + lm_head + argmax) timed separately.  Returns
    (cpu_baseline dict, measured seconds of one sample step)."""
    import torch
    from oracle.opt_ref import CpuPolicy1Runner
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    h, H, f, L = cfg.hidden_size, cfg.num_attention_heads, cfg.ffn_dim, cfg.num_hidden_layers
    B, S, new = args.batch_size, args.input_tokens, args.max_new_tokens
    # probe the host's bf16 GEMM rate to size the sample
    a = torch.randn(512, h).to(torch.bfloat16)
    w = torch.randn(h, h).to(torch.bfloat16)
    torch.nn.functional.linear(a, w)
    t0 = time.perf_counter()
    torch.nn.functional.linear(a, w)
    tf = 2 * 512 * h * h / (time.perf_counter() - t0) / 1e12
    per_step = budget_s / max(1, steps + warmup) if steps + warmup > 1 else budget_s
    per_step = max(per_step, 3.0)
    layer_flops = 24.0 * h * h * B * S + 4.0 * S * S * h * B / 2
    layer_bytes = 2.0 * 12 * h * h
    est_layer = layer_flops / (tf * 1e12) + (new - 1) * layer_bytes / 60e9      # prefill at the probed rate, decode at ~60 GB/s
    whole = est_layer * L <= per_step
    n_layers = L if whole else min(2, L)
    Bs = B
    while not whole and Bs > 1 and est_layer * n_layers * (Bs / B) > per_step:
        Bs //= 2
    r = CpuPolicy1Runner(h, H, f, n_layers, Bs, S + new)
    times = []
    for i in range(warmup + steps):
        t = r.run(S, new)
        if i >= warmup:
            times.append(t)
    t_sample = sum(times) / len(times)
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
    t_total = t_sample * (L / n_layers) + t_head * new
    tok_s = (Bs * new) / t_total
    if whole:
        sample = (f"the WHOLE {cfg.name} stack ({L} layers) at batch {Bs}, input {S}, {new} new tokens -- nothing extrapolated -- plus "
                  f"{new} x (final LayerNorm + lm_head + argmax) timed separately ({t_head * 1e3:.1f} ms each); embedding lookups "
                  f"excluded; host GEMM probe {tf:.2f} TFLOP/s")
    else:
        sample = (f"{n_layers} consecutive of {L} decoder layers of {cfg.name} at batch {Bs} (of {B}), input {S}, {new} new tokens; "
                  f"stack time = sample time x {L}/{n_layers}, plus {new} x (final LayerNorm + lm_head + argmax) timed separately "
                  f"({t_head * 1e3:.1f} ms each); embedding lookups excluded; host GEMM probe {tf:.2f} TFLOP/s")
    return {"value": tok_s, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample,
            "sample_seconds": t_sample, "sample_layers": n_layers, "sample_batch": Bs, "head_seconds": t_head,
            "extrapolated_full_step_seconds": t_total * (B / Bs)}, t_sample


def _ulp(x):
    import torch
    return 2.0 ** (torch.floor(torch.log2(x.abs().clamp_min(1e-30))) - 7)


def parity_check(m, cfg, args, st, ids_dev, out_tokens, rank, world, dev):
    """The timed model's outputs against the oracle (oracle/opt_ref.py: the reference's eager op sequence on the
    unsharded weights, full depth, same GPU), OUTSIDE the timed region.  Rank 0 computes; returns the dict for the
    JSON line (None on other ranks).  Checked: last-layer hidden state of every prompt position after the prefill
    (max relative error), the greedy tokens of the last timed generate(), and -- per sequence -- the oracle's own top-2
    logit margin at the first step where the two token streams part (a near-tie there means fp32 summation order, not
    an error, decided the token; after it the contexts differ, so later tokens are not comparable)."""
    import torch
    from oracle import opt_ref
    from lia_b200.weights import random_embeddings, random_layer
    if rank != 0:
        return None
    dec = m.model.decoder
    L, h, f = cfg.num_hidden_layers, cfg.hidden_size, cfg.ffn_dim
    B, S, new = args.batch_size, args.input_tokens, args.max_new_tokens
    if dec.layout.dp != dec.layout.d:
        return {"skipped": "head-padded layout: oracle weights are not views of the slabs"}
    if dev.type == "cuda":
        torch.cuda.empty_cache()
        free = torch.cuda.mem_get_info(dev)[0]
    else:
        free = 1 << 42
    Wl = 2.0 * (12 * h * h + 13 * h)
    cache_bytes = L * 4.0 * (S + new) * B * h
    scratch = 6.0 * B * S * max(f, S * cfg.num_attention_heads) * 2 + (8 << 30)
    views_ok = world == 1 and dec.n_resident == L

    def split(v):
        hq = h
        w = {k: v[k] for k in ("ln1_w", "ln1_b", "o_w", "o_b", "ln2_w", "ln2_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b")}
        w["q_w"], w["k_w"], w["v_w"] = v["qkv_w"][:hq], v["qkv_w"][hq:2 * hq], v["qkv_w"][2 * hq:]
        w["q_b"], w["k_b"], w["v_b"] = v["qkv_b"][:hq], v["qkv_b"][hq:2 * hq], v["qkv_b"][2 * hq:]
        return w

    def gen(i):      # the unsharded layer exactly as init_weights() drew it (same per-layer seed on every rank)
        return random_layer(h, f, 0 * 100003 + 1000 + i, dev, args.weights, cfg.init_std, 0.0, 0.0)

    class Lazy:      # layers regenerated on every pass (models whose full weights do not fit next to the shard)
        def __len__(self):
            return L

        def __iter__(self):
            for i in range(L):
                yield gen(i)

    if views_ok:
        layers, mode = [split(v) for v in dec.resident_views], "oracle reads the model's own resident slabs"
    elif free > L * Wl + cache_bytes + scratch:
        layers, mode = [gen(i) for i in range(L)], "unsharded weights regenerated from the per-layer seeds on rank 0"
    else:
        layers, mode = Lazy(), "unsharded weights regenerated layer by layer (too large to hold): prefill and first token only"
    full = not isinstance(layers, Lazy) and free > (0 if views_ok else L * Wl) + cache_bytes + scratch
    om = {"H": cfg.num_attention_heads, "layers": layers, "pre_ln": cfg.do_layer_norm_before, "embed_tokens": dec.embed_tokens,
          "embed_positions": dec.embed_positions, "final_ln_w": dec.final_ln_w, "final_ln_b": dec.final_ln_b,
          "project_in": dec.project_in, "project_out": dec.project_out}

    mb = max(1, B // max(1, args.num_minibatch))
    per_layer = world == 1 and dec.n_resident == L and dev.type == "cuda"

    class Last:
        """Receives every layer's hidden state from the oracle; keeps only the last one (48 x 235 MB otherwise).  At one GPU it
        also runs THIS build's layer on the oracle's own input of that layer (first minibatch) -- the north-star criterion is a
        per-layer error, which a chained comparison over 48 layers cannot show (rounding noise accumulates like a random walk)."""

        def __init__(self, x0):
            self.x, self.n, self.errs = x0, 0, []

        def append(self, x):
            if per_layer:
                H_, d_ = cfg.num_attention_heads, h // cfg.num_attention_heads
                rows = self.x[:mb].reshape(mb * S, h).clone()
                kc = torch.zeros(S, mb, H_, d_, dtype=torch.bfloat16, device=dev)
                vc = torch.zeros_like(kc)
                dec.layer_rows(dec.resident_views[self.n], rows, kc, vc, mb, S, 0, 0, st.ws)
                ref_rows = x[:mb].reshape(mb * S, h).float()
                self.errs.append(((rows.float() - ref_rows).abs().max() / ref_rows.abs().max()).item())
            self.x, self.n = x, self.n + 1

    res = {"oracle": "oracle/opt_ref.py (reference op sequence, torch eager on the same GPU, full depth, unsharded)", "weights": mode}
    with torch.no_grad():
        mask = torch.ones(B, S, dtype=torch.long, device=dev)
        n_steps = new if full else 1
        Hh = cfg.num_attention_heads
        cache = [(torch.zeros(S + n_steps, B, Hh, h // Hh, dtype=torch.bfloat16, device=dev),
                  torch.zeros(S + n_steps, B, Hh, h // Hh, dtype=torch.bfloat16, device=dev)) for _ in range(L)]   # A:471-472
        last = Last(opt_ref.embed(om, ids_dev, mask, 0) if per_layer else None)
        hid = opt_ref.decoder_forward(om, ids_dev, mask, cache, 0, collect=last)
        if per_layer:
            res["per_layer_rel_err_max"] = max(last.errs)
            res["per_layer_rel_err_mean"] = sum(last.errs) / len(last.errs)
            res["per_layer_rel_err_over_1e-2"] = sum(1 for e in last.errs if e > 1e-2)
            res["per_layer_note"] = (f"each of the {L} layers run by this build on the ORACLE's input of that layer (first minibatch, {mb} x {S} rows): "
                                     "the north-star bound is 1e-2 per layer")
        ours = st.x.view(B, S, h).float()
        ref = last.x.float()
        res["prefill_hidden_rel_err"] = ((ours - ref).abs().max() / ref.abs().max()).item()   # CHAINED over all layers
        res["prefill_hidden_rows"] = B * S
        del ours, ref, last
        # greedy tokens: free-running oracle vs the tokens the timed generate() returned
        toks = out_tokens.to(dev)
        ref_ids = ids_dev
        logits_t = []
        cur, past = None, S
        eos = cfg.eos_token_id
        for t in range(n_steps):
            if t > 0:
                mask = torch.cat([mask, mask.new_ones(B, 1)], dim=-1)
                hid = opt_ref.decoder_forward(om, cur, mask, cache, past)
                past += 1
            lg = opt_ref.lm_logits(om, hid)[:, -1, :].float()
            lg[:, eos] = -float("inf")
            logits_t.append(lg)
            cur = lg.argmax(-1)[:, None]
            ref_ids = torch.cat([ref_ids, cur], dim=-1)
        eq = (toks[:, S:S + n_steps] == ref_ids[:, S:])
        res["tokens_compared"] = int(eq.numel())
        res["tokens_equal_frac"] = eq.float().mean().item()
        res["sequences_identical"] = int(eq.all(dim=1).sum())
        res["sequences"] = B
        worst, n_div = 0.0, 0
        for b in range(B):
            bad = (~eq[b]).nonzero()
            if bad.numel():
                t = int(bad[0])
                n_div += 1
                lg = logits_t[t][b]
                margin = (lg[ref_ids[b, S + t]] - lg[toks[b, S + t]]).item()
                worst = max(worst, margin / _ulp(lg[ref_ids[b, S + t]]).item())
        res["first_divergences"] = n_div
        res["worst_first_divergence_margin_bf16_ulps"] = worst
        res["note"] = ("free-running comparison: after a sequence's first divergence its context differs, so only the tokens up to "
                       "it count; every first divergence is classified by the oracle's own top-2 margin there (<= 2 ulp = a tie "
                       "that fp32 summation order decides)")
    del cache, layers, om
    if dev.type == "cuda":
        torch.cuda.empty_cache()
    return res


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
                + (f" [DEBUG depth {args.layers}: not a BASELINE config]" if args.layers else
                   f" ({CONFIG_LABEL[args.config_name]})" if args.config_name else ""))
    config = {"workload": workload, "parallelism": f"tp{world}", "l2": "inputs_exceed_l2 (weights+KV per step >> 126 MB)",
              "cuda_graphs": not args.no_graphs}
    if args.quick:
        config["quick"] = "ONE warm-up pass, one timed pass with host buffers (value == e2e): a minutes-long PCIe-bound generate()"

    if args.impl == "reference":
        if rank != 0:
            return 0
        cb, t_step = cpu_baseline(args, cfg, steps=max(1, args.steps), warmup=min(args.warmup, 1), budget_s=150.0)
        line = {"impl": "reference", "metric": "tokens/s", "value": cb["value"], "unit": "tokens/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3,
                "ms_per_step_note": "measured time of ONE sample step (cpu_baseline.sample); the full-workload step it implies is "
                                    "cpu_baseline.extrapolated_full_step_seconds",
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
    n_warm = 1 if args.quick else max(args.warmup, 3)
    for _ in range(n_warm - 1):
        m.generate(ids_dev, **kw)
    dec = m.model.decoder
    stream0 = dec.streamer.stats() if dec.streamer is not None else None

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
        sec, out, prefill, decode = timed(ids_host if args.quick else ids_dev)
    stream1 = dec.streamer.stats() if dec.streamer is not None else None
    if args.quick:      # one timed pass only, through the public call with host buffers: `value` and `e2e` are the same measurement
        sec_e2e, out_h = sec, out.cpu()
    else:
        sec_e2e, out_h, _, _ = timed(ids_host)
    assert out_h.device.type == "cpu" and out_h.shape == (B, S + new)
    tok = B * new * args.steps
    value, e2e = tok / sec, tok / sec_e2e
    t_prefill = sum(prefill) / len(prefill)
    t_decode = sum(decode) / max(1, len(decode))
    st = next(iter(m._states.values()))

    # ---- parity of the timed result (outside the timed region; st.x still holds the last prefill's hidden states)
    parity = None
    if args.weights == "dummy":
        parity = {"skipped": "dummy U[0,1) weights (utils/opt-weight-gen.py:61-62) overflow bf16 within a layer: throughput-only workload"}
    elif not args.no_parity:
        try:
            parity = parity_check(m, cfg, args, st, ids_dev, out, rank, world, dev)
        except torch.OutOfMemoryError as e:      # the checker must never cost the line its throughput numbers
            parity = {"skipped": f"oracle ran out of memory: {str(e)[:120]}"}
            if dev.type == "cuda":
                torch.cuda.empty_cache()
        tp.barrier()

    # ---- roofline of the dominant kernel: instrument one prefill pass, CUDA events around every GEMM launch
    L, h, f, V = cfg.num_hidden_layers, cfg.hidden_size, cfg.ffn_dim, cfg.vocab_size
    recs = []
    orig, orig_ar = ops.gemm, ops.gemm_allreduce

    def timed_gemm(a, w, *a2, **k2):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        r = orig(a, w, *a2, **k2)
        e.record()
        recs.append((a.shape[0], w.shape[0], a.shape[1], s, e, "gemm"))
        return r

    def timed_gemm_ar(a, w, *a2, **k2):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        r = orig_ar(a, w, *a2, **k2)
        e.record()
        recs.append((a.shape[0], w.shape[0], a.shape[1], s, e, "gemm_allreduce"))
        return r

    ops.gemm, ops.gemm_allreduce = timed_gemm, timed_gemm_ar
    tp.barrier()
    m._prefill(st, args.num_minibatch, -1)
    ops.gemm, ops.gemm_allreduce = orig, orig_ar
    torch.cuda.synchronize(dev)
    big = [(M_, N_, K_, s.elapsed_time(e), kind) for (M_, N_, K_, s, e, kind) in recs if M_ > 128]
    roof = None
    if big:
        flops = sum(2.0 * M_ * N_ * K_ for M_, N_, K_, _, _ in big)
        ms = sum(t for *_, t, _ in big)
        ach = flops / (ms / 1e3) / 1e12
        traffic, traffic_note = None, None
        if not args.layers and world == 1:
            traffic, traffic_note = committed_traffic(os.environ.get("LIA_GEMM_2CTA", "1") != "0")
        roof = {"kernel": "lia_gemm_tcgen05_kernel (prefill projections" + (", incl. the fused row-parallel GEMM + all-reduce launches)" if world > 1 else ")"),
                "bound": "tensor", "achieved": ach,
                "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": ach / pk["bf16_sustained"], "traffic": traffic,
                "frac_of_burst_peak": ach / pk["bf16_burst"],
                "launches": len(big), "avg_launch_ms": ms / len(big), "flops_per_launch": flops / len(big),
                "peak_source": pk["source"] + ", sustained figure (kernel timed inside a long step); burst "
                + f"{pk['bf16_burst']}", "share_of_step": (ms / 1e3) / (sec / args.steps),
                "traffic_note": traffic_note}
        if world > 1:
            ar = [(M_, N_, K_, t) for (M_, N_, K_, t, kind) in big if kind == "gemm_allreduce"]
            pl = [(M_, N_, K_, t) for (M_, N_, K_, t, kind) in big if kind == "gemm"]
            if ar and pl:
                roof["column_parallel_tflops"] = sum(2.0 * a * b * c for a, b, c, _ in pl) / (sum(t for *_, t in pl) / 1e3) / 1e12
                roof["row_parallel_fused_allreduce_tflops"] = sum(2.0 * a * b * c for a, b, c, _ in ar) / (sum(t for *_, t in ar) / 1e3) / 1e12
                roof["row_parallel_fused_allreduce_ms_per_prefill"] = sum(t for *_, t in ar)
                roof["exchange_bytes_per_prefill_per_rank"] = sum(2.0 * a * b * 2 * (world - 1) / world for a, b, _, _ in ar)
    # decode step: algorithmic bytes = L*W_l + L*4*B*T*h + 2*h*V   (SURVEY.md 8d), per rank
    Wl = 2.0 * (12 * h * h + 13 * h)
    Tavg = S + (new - 1) / 2.0 + 0.5
    dec_bytes = (L * Wl + L * 4.0 * B * Tavg * h) / world + 2.0 * h * V
    streamed = dec.streamer is not None or st.spill is not None
    roof_dec = {"kernel": "decode step (swap-AB tcgen05 GEMMs + flash-decoding attention + LN" + (", CUDA graph)" if not streamed and not args.no_graphs else ", eager: layers/KV streamed from pinned host)"),
                "bound": "hbm" if not streamed else "pcie",
                "achieved": dec_bytes / t_decode / 1e9 if t_decode > 0 else None, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": (dec_bytes / t_decode / 1e9) / pk["hbm_gbs"] if t_decode > 0 else None, "traffic": None,
                "bytes_per_step": dec_bytes, "ms_per_decode_step": t_decode * 1e3, "peak_source": pk["source"]}

    if rank == 0:
        line = {"metric": "tokens/s", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
                "warmup": n_warm, "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
                "clocks": clk.summary(),
                "e2e": {"value": e2e, "unit": "tokens/s", "h2d_bytes_per_step": B * S * 8, "d2h_bytes_per_step": B * (S + new) * 8},
                "gpu_launches": launches_per_step * args.steps,
                "prefill_ms": t_prefill * 1e3, "decode_ms_per_step": t_decode * 1e3,
                "roofline": roof, "roofline_decode": roof_dec, "parity": parity}
        if stream1 is not None:
            db, dms = stream1["bytes"] - stream0["bytes"], stream1["copy_ms"] - stream0["copy_ms"]
            line["weight_stream"] = {"resident_layers": dec.n_resident, "streamed_layers": L - dec.n_resident,
                                     "h2d_gb_per_step": db / 1e9 / args.steps, "pcie_gbs_while_copying": (db / 1e9) / (dms / 1e3) if dms > 0 else None,
                                     "pcie_gbs_over_step": (db / 1e9) / sec, "host_layer_pool": dec.host_pool or None}
        if st.spill is not None:
            sp = st.spill.stats()
            line["kv_spill"] = {"spilled_layers": sp["layers"], "resident_layers": st.kv_resident, "host_gb": sp["host_bytes"] / 1e9}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"], _ = cpu_baseline(args, cfg)
        print(json.dumps(line), file=json_out)
        json_out.flush()
    tp.barrier()
    return 0


if __name__ == "__main__":
    sys.exit(main())
