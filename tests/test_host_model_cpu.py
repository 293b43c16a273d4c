"""HOST logic of the model surface on CPU: ``modeling_opt`` driven through a stock-PyTorch stand-in for the kernel layer
(tests/cpu_ops_emulation.py -- test infrastructure) and compared BIT FOR BIT with the oracle and with the goldens made
from the reference's own layer code.  What this pins without a GPU: the order and operands of the seven launches of a
layer, the minibatch loop and its cache windows, the KV-cache / past_key_values bookkeeping, learned positions under
padding, the last-token head, eos suppression and the token matrix of the greedy loop, and the three reference faces
(layer, decoder, model).  The kernels themselves are checked on the GPU (tests/test_gpu_*.py)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lia_b200  # noqa: E402
from oracle import opt_ref  # noqa: E402
from cpu_ops_emulation import cpu_ops  # noqa: E402,F401  (fixture)
from test_oracle_golden import LAYER_CASES, load_layer_case  # noqa: E402

BF16 = torch.bfloat16


def oracle_model(m):
    """Oracle model dict over the SAME weights as the product model ``m`` (unfused q/k/v views of its slabs)."""
    dec = m.model.decoder
    hq = dec.layout.hq
    layers = []
    for v in dec.resident_views:
        w = {k: v[k] for k in ("ln1_w", "ln1_b", "o_w", "o_b", "ln2_w", "ln2_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b")}
        w["q_w"], w["k_w"], w["v_w"] = v["qkv_w"][:hq], v["qkv_w"][hq:2 * hq], v["qkv_w"][2 * hq:]
        w["q_b"], w["k_b"], w["v_b"] = v["qkv_b"][:hq], v["qkv_b"][hq:2 * hq], v["qkv_b"][2 * hq:]
        layers.append(w)
    om = {"H": m.config.num_attention_heads, "layers": layers, "embed_tokens": dec.embed_tokens,
          "embed_positions": dec.embed_positions, "final_ln_w": dec.final_ln_w, "final_ln_b": dec.final_ln_b,
          "pre_ln": m.config.do_layer_norm_before}
    for k in ("project_in", "project_out"):
        om[k] = getattr(dec, k, None)
    return om


def tiny(**kw):
    base = dict(hidden_size=128, num_hidden_layers=3, num_attention_heads=2, ffn_dim=512, vocab_size=384, max_position_embeddings=64)
    base.update(kw)
    cfg = lia_b200.OPTConfig(**base)
    m = lia_b200.OPTForCausalLM(cfg, "cpu").init_weights(seed=4, bias_std=0.05, ln_std=0.1)
    m.use_cuda_graphs = False
    return cfg, m


@pytest.mark.parametrize("num_minibatch", [1, 2, 3])
def test_generate_matches_oracle_bit_for_bit(cpu_ops, num_minibatch):
    cfg, m = tiny()
    B, S, new = 6, 9, 5
    ids = torch.randint(3, cfg.vocab_size, (B, S), generator=torch.Generator().manual_seed(1))
    toks = m.generate(ids, max_new_tokens=new, min_new_tokens=new, prefill_policy=0, decoding_policy=0, num_minibatch=num_minibatch)
    with torch.no_grad():
        ref = opt_ref.greedy_generate(oracle_model(m), ids, new)
    assert torch.equal(toks, ref)
    # cache rows of every layer equal the oracle's (time-major [T, B, H, d], A:471-472)
    st = next(iter(m._states.values()))
    cache = opt_ref.new_cache(oracle_model(m), B, S + new)
    with torch.no_grad():
        hid = opt_ref.decoder_forward(oracle_model(m), ids, torch.ones(B, S, dtype=torch.long), cache, 0)
    for li in range(cfg.num_hidden_layers):
        assert torch.equal(st.kc[li][:S], cache[li][0][:S]) and torch.equal(st.vc[li][:S], cache[li][1][:S])
    # launch sequence of one layer in prefill: LN, fused QKV, attention, out_proj(+res), LN, fc1(+relu), fc2(+res)
    names = [c for c in cpu_ops.log if c[0] != "embed"][:7]
    assert [c[0] for c in names] == ["layernorm", "gemm", "attn_prefill", "gemm", "layernorm", "gemm", "gemm"]
    assert [c[1] for c in names if c[0] == "gemm"] == [3, 2, 1, 2]          # QKV, BIAS_RESIDUAL, BIAS_RELU, BIAS_RESIDUAL
    assert hid.shape == (B, S, cfg.hidden_size)


def test_generate_padded_prompt_and_eos_suppression(cpu_ops):
    cfg, m = tiny()
    B, S, new = 4, 8, 4
    ids = torch.randint(3, cfg.vocab_size, (B, S), generator=torch.Generator().manual_seed(2))
    ids[0, :3] = cfg.pad_token_id          # left padding moves the learned positions (M:368-378)
    ids[1, -2:] = cfg.pad_token_id
    toks = m.generate(ids, max_new_tokens=new, min_new_tokens=new)
    with torch.no_grad():
        ref = opt_ref.greedy_generate(oracle_model(m), ids, new)
    assert torch.equal(toks, ref)
    assert not (toks[:, S:] == cfg.eos_token_id).any()
    # an explicit mask takes precedence over pad ids
    mask = torch.ones(B, S, dtype=torch.long)
    mask[2, :2] = 0
    toks2 = m.generate(ids, max_new_tokens=new, min_new_tokens=new, attention_mask=mask)
    with torch.no_grad():
        ref2 = opt_ref.greedy_generate(oracle_model(m), ids, new, attention_mask=mask)
    assert torch.equal(toks2, ref2)
    # token_latency: (ids, per-token latency list) as greedy_search.py:455-456
    m.config.token_latency = True
    out = m.generate(ids, max_new_tokens=new, min_new_tokens=new)
    assert isinstance(out, tuple) and torch.equal(out[0], toks) and len(out[1]) == new


def test_generate_edge_shapes(cpu_ops):
    """One new token (prefill only), a batch the minibatch count does not divide (M:1178 floors; the remainder forms a
    last short minibatch), more minibatches than sequences, and the length checks."""
    cfg, m = tiny()
    om = oracle_model(m)
    for B, S, new, nmb in [(2, 5, 1, 1), (5, 6, 3, 2), (1, 4, 2, 4), (7, 3, 2, 3)]:
        ids = torch.randint(3, cfg.vocab_size, (B, S), generator=torch.Generator().manual_seed(B * 10 + S))
        toks = m.generate(ids, max_new_tokens=new, min_new_tokens=new, num_minibatch=nmb)
        with torch.no_grad():
            ref = opt_ref.greedy_generate(om, ids, new)
        assert toks.shape == (B, S + new) and torch.equal(toks, ref), (B, S, new, nmb)
    with pytest.raises(ValueError, match="max_new_tokens"):
        m.generate(ids, max_new_tokens=0)
    with pytest.raises(ValueError, match="max_position_embeddings"):
        m.generate(ids, max_new_tokens=cfg.max_position_embeddings)
    with pytest.raises(NotImplementedError):
        m.generate(ids, max_new_tokens=2, num_beams=4)
    with pytest.raises(ValueError, match="attention_mask"):
        m.generate(ids, max_new_tokens=2, attention_mask=torch.ones(B, S + 1, dtype=torch.long))


def test_forward_face_prefill_then_decode(cpu_ops):
    """models.py:371-445: logits [B,1,V] of the last position + the 4-tuple cache whose marker carries the length."""
    cfg, m = tiny()
    om = oracle_model(m)
    B, S, new = 3, 7, 3
    ids = torch.randint(3, cfg.vocab_size, (B, S), generator=torch.Generator().manual_seed(5))
    mask = torch.ones(B, S, dtype=torch.long)
    cache = opt_ref.new_cache(om, B, S + new)
    logits, past = m(input_ids=ids, attention_mask=mask, prefill_policy=0, decoding_policy=0, num_minibatch=1, max_new_tokens=new)
    with torch.no_grad():
        ref = opt_ref.lm_logits(om, opt_ref.decoder_forward(om, ids, mask, cache, 0))
    assert logits.shape == (B, 1, cfg.vocab_size) and torch.equal(logits, ref)
    assert len(past) == cfg.num_hidden_layers and past[0][0].shape[2] == S and past[0][1].shape == (S + new, B, 2, 64)
    cur = torch.argmax(logits[:, -1].float(), dim=-1)[:, None]
    for t in range(new - 1):
        mask = torch.cat([mask, mask.new_ones(B, 1)], dim=-1)
        logits, past = m(input_ids=cur, attention_mask=mask, past_key_values=past, max_new_tokens=new)
        with torch.no_grad():
            ref = opt_ref.lm_logits(om, opt_ref.decoder_forward(om, cur, mask, cache, S + t))
        assert torch.equal(logits, ref), t
        assert past[0][0].shape[2] == S + t + 1
        cur = torch.argmax(logits[:, -1].float(), dim=-1)[:, None]
    with pytest.raises(ValueError):
        m(input_ids=cur, attention_mask=mask[:, :-1], past_key_values=past)       # M:1127-1131
    with pytest.raises(NotImplementedError):
        m(input_ids=ids, prefill_policy=1)                                        # full-CPU policy is not in this build


@pytest.mark.parametrize("name", LAYER_CASES)
def test_layer_face_vs_reference_goldens(cpu_ops, golden_dir, name):
    """decoder.py:172-183, 323-335 through OPTDecoderLayer.forward with the reference's 16-entry ``gpu_layer`` list:
    outputs and cache rows bit-identical to what the reference's own layer code produced (oracle/gen_golden.py)."""
    meta, w, xs, ys, kc_ref, vc_ref = load_layer_case(golden_dir, name)
    B, S, h, H, new = (meta[k] for k in ("B", "S", "h", "H", "new"))
    cfg = lia_b200.OPTConfig(hidden_size=h, num_hidden_layers=1, num_attention_heads=H, ffn_dim=4 * h, vocab_size=64,
                             max_position_embeddings=64, do_layer_norm_before=meta["pre_ln"])
    m = lia_b200.OPTForCausalLM(cfg, "cpu")
    layer = m.model.decoder.layers[0]
    gpu_layer = [w[k] for k in lia_b200.weights.LAYER_KEYS]
    past = None
    for step, (x, y_ref) in enumerate(zip(xs, ys)):
        out = layer(x, past_key_value=past, use_cache=True, gpu_layer=gpu_layer, policy=0, max_new_tokens=new)
        y, past = out[0], out[1]
        assert torch.equal(y.view(torch.int16), y_ref.view(torch.int16)), (name, step)
        assert past[0].shape[2] == S + step
        k_new, v_new = out[2], out[3]                                # policy 0 also returns this call's K/V rows (D:331-333)
        lo = 0 if step == 0 else S + step - 1
        assert torch.equal(k_new, kc_ref[lo:S + step]) and torch.equal(v_new, vc_ref[lo:S + step])
    assert torch.equal(past[1][:S + new], kc_ref) and torch.equal(past[2][:S + new], vc_ref)


# ------------------------------------------------------------------ opt-350m's shape: post-LN, project_in/out, no final LN

def test_postln_projected_model_generate_matches_oracle(cpu_ops):
    cfg, m = tiny(do_layer_norm_before=False, word_embed_proj_dim=64)
    dec = m.model.decoder
    assert dec.final_ln_w is None and dec.project_in.shape == (128, 64) and dec.project_out.shape == (64, 128)
    assert dec.embed_tokens.shape == (cfg.vocab_size, 64)
    B, S, new = 5, 8, 4
    ids = torch.randint(3, cfg.vocab_size, (B, S), generator=torch.Generator().manual_seed(7))
    ids[0, :2] = cfg.pad_token_id
    toks = m.generate(ids, max_new_tokens=new, min_new_tokens=new, num_minibatch=2)
    with torch.no_grad():
        ref = opt_ref.greedy_generate(oracle_model(m), ids, new)
    assert torch.equal(toks, ref)
    # launch sequence of one post-LN layer: fused QKV, attention, out_proj(+res), LN, fc1(+relu), fc2(+res), LN
    seq = [c for c in cpu_ops.log]
    first_qkv = next(i for i, c in enumerate(seq) if c[0] == "gemm" and c[1] == 3)
    assert [c[0] for c in seq[first_qkv:first_qkv + 7]] == ["gemm", "attn_prefill", "gemm", "layernorm", "gemm", "gemm", "layernorm"]
    # embeddings: token rows, position rows, then project_in with the positions as its residual (M:1139-1142)
    assert [c[0] for c in seq[:3]] == ["embed", "embed", "gemm"] and seq[2][1] == 2 and seq[2][3:] == (128, 64)


def test_postln_projected_model_matches_stock_transformers(cpu_ops, golden_dir):
    """The whole surface on an HF state dict of opt-350m's architecture (tests/golden/model_hf_postln_tiny.npz, made by
    stock transformers in fp32): same greedy tokens, bf16 logits within the reference's nightly tolerance (0.1)."""
    z = np.load(os.path.join(golden_dir, "model_hf_postln_tiny.npz"))
    bf = lambda a: torch.from_numpy(a.view(np.int16).copy()).view(BF16)  # noqa: E731
    sd = {k[3:]: bf(z[k]) for k in z.files if k.startswith("sd:")}
    cfg = lia_b200.OPTConfig(hidden_size=int(z["h"]), num_hidden_layers=int(z["L"]), num_attention_heads=int(z["H"]),
                             ffn_dim=4 * int(z["h"]), vocab_size=int(z["V"]), max_position_embeddings=int(z["P"]),
                             do_layer_norm_before=bool(int(z["pre_ln"])), word_embed_proj_dim=int(z["word_dim"]))
    m = lia_b200.OPTForCausalLM(cfg, "cpu").load_state_dict(sd)
    m.use_cuda_graphs = False
    ids = torch.from_numpy(z["input_ids"])
    logits, _ = m(input_ids=ids, attention_mask=torch.ones_like(ids), max_new_tokens=int(z["new"]))
    assert np.abs(logits[:, -1].float().numpy() - z["prefill_last_logits"]).max() < 0.1
    toks = m.generate(ids, max_new_tokens=int(z["new"]), min_new_tokens=int(z["new"]))
    with torch.no_grad():
        ref = opt_ref.greedy_generate(oracle_model(m), ids, int(z["new"]))
    assert torch.equal(toks, ref)
    # bf16 rounding may flip a near-tie against the fp32 golden tokens; the first token of every row is far from one here
    assert np.array_equal(toks[:, ids.shape[1]].numpy(), z["tokens"][:, ids.shape[1]])


def test_mismatched_embeddings_are_refused(cpu_ops):
    cfg = lia_b200.OPTConfig(hidden_size=128, num_hidden_layers=1, num_attention_heads=2, ffn_dim=512, vocab_size=96,
                             max_position_embeddings=32, do_layer_norm_before=False, word_embed_proj_dim=64)
    m = lia_b200.OPTForCausalLM(cfg, "cpu")
    e = lia_b200.weights.random_embeddings(96, 128, 32, 0, embed_dim=64, final_ln=True)       # a final LN a post-LN model lacks
    with pytest.raises(ValueError, match="final LayerNorm"):
        m.model.decoder.load_embeddings(e)
    e = lia_b200.weights.random_embeddings(96, 128, 32, 0, embed_dim=128, final_ln=False)     # no projections, wrong width
    with pytest.raises(ValueError, match="project_in"):
        m.model.decoder.load_embeddings(e)


# ------------------------------------------------------------------ head_dim other than 64 / 128 (opt-2.7b: 80): zero-padded heads

def _unpadded_oracle_model(cfg, seed, bias_std, ln_std):
    """Oracle dict from the SAME generators init_weights() uses, but unpadded (the product's slabs hold padded heads)."""
    from lia_b200.weights import random_embeddings, random_layer
    e = random_embeddings(cfg.vocab_size, cfg.hidden_size, cfg.max_position_embeddings, seed * 100003 + 17, "cpu", "normal",
                          cfg.init_std, ln_std, cfg.pad_token_id, embed_dim=cfg.embed_dim, final_ln=cfg.do_layer_norm_before)
    layers = [random_layer(cfg.hidden_size, cfg.ffn_dim, seed * 100003 + 1000 + i, "cpu", "normal", cfg.init_std, bias_std, ln_std)
              for i in range(cfg.num_hidden_layers)]
    return {"H": cfg.num_attention_heads, "layers": layers, "pre_ln": cfg.do_layer_norm_before, **e}


@pytest.mark.parametrize("d", [80, 32, 96])
def test_padded_head_dim_matches_unpadded_oracle(cpu_ops, d):
    H = 4
    cfg = lia_b200.OPTConfig(hidden_size=H * d, num_hidden_layers=2, num_attention_heads=H, ffn_dim=4 * H * d, vocab_size=384,
                             max_position_embeddings=64)
    m = lia_b200.OPTForCausalLM(cfg, "cpu").init_weights(seed=4, bias_std=0.05, ln_std=0.1)
    m.use_cuda_graphs = False
    lay = m.layout
    dp = 64 if d <= 64 else 128
    assert (lay.d, lay.dp, lay.hq) == (d, dp, H * dp)
    v = m.model.decoder.resident_views[0]
    assert v["qkv_w"].shape == (3 * H * dp, H * d) and v["o_w"].shape == (H * d, H * dp)
    q = v["qkv_w"][:H * dp].view(H, dp, H * d)
    assert not q[:, d:].any() and not v["qkv_b"].view(3, H, dp)[:, :, d:].any() and not v["o_w"].view(H * d, H, dp)[:, :, d:].any()
    om = _unpadded_oracle_model(cfg, 4, 0.05, 0.1)
    assert torch.equal(q[:, :d].reshape(H * d, H * d), om["layers"][0]["q_w"])
    B, S, new = 3, 9, 4
    ids = torch.randint(3, cfg.vocab_size, (B, S), generator=torch.Generator().manual_seed(1))
    ones = torch.ones(B, S, dtype=torch.long)
    hidden, past = m.model.decoder(input_ids=ids, attention_mask=ones, max_new_tokens=new)
    cache = opt_ref.new_cache(om, B, S + new)
    with torch.no_grad():
        href = opt_ref.decoder_forward(om, ids, ones, cache, 0)
    # padding only adds exact zeros to every dot product: the values are the unpadded ones up to fp32 summation order
    err = ((hidden.float() - href.float()).abs().max() / href.float().abs().max()).item()
    assert err <= 1e-2, err
    for li in range(cfg.num_hidden_layers):
        k, kref = past[li][1], cache[li][0]
        assert k.shape == (S + new, B, H, dp) and not k[..., d:].any() and not past[li][2][..., d:].any()
        if li == 0:                       # K/V of the first layer: same products, same order -> identical bits
            assert torch.equal(k[:S, ..., :d], kref[:S])
    toks = m.generate(ids, max_new_tokens=new, min_new_tokens=new, num_minibatch=2)
    with torch.no_grad():
        lg = []
        ref = opt_ref.greedy_generate(om, ids, new, collect_logits=lg)
    l0 = lg[0].clone()
    l0[:, cfg.eos_token_id] = -float("inf")
    top2 = l0.topk(2, dim=-1).values
    safe = (top2[:, 0] - top2[:, 1]) > 8 * 2.0 ** -8 * top2[:, 0].abs()
    assert torch.equal(toks[safe, S], ref[safe, S])
    with pytest.raises(NotImplementedError, match="head_dim"):
        lia_b200.OPTForCausalLM(lia_b200.OPTConfig(hidden_size=4 * 136, num_attention_heads=4, ffn_dim=64), "cpu")


# ------------------------------------------------------------------ bench.py's own arm, dry-run

def test_bench_line_contract_dry_run(cpu_ops, monkeypatch):
    """bench.py's main body executed on the kernel stand-in (device swapped to CPU, event times faked): no number here
    means anything, but every key of the driver's JSON contract must be present and well-formed, and the instrumented
    prefill pass must see exactly the projection GEMMs (4 per layer per minibatch)."""
    import io
    import json
    import types

    class Ev:
        def __init__(self, enable_timing=False):
            pass

        def record(self, *a):
            pass

        def elapsed_time(self, other):
            return 1.0
    monkeypatch.setattr(torch.cuda, "Event", Ev)
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a, **k: None)
    src = open(os.path.join(ROOT, "bench.py")).read()
    for old, new in (('torch.device("cuda", local)', 'torch.device("cpu")'), (".pin_memory()", ""),
                     ("m.use_cuda_graphs = not args.no_graphs", "m.use_cuda_graphs = False")):
        assert old in src
        src = src.replace(old, new)
    mod = types.ModuleType("bench_dry")
    mod.__dict__["__file__"] = os.path.join(ROOT, "bench.py")
    exec(compile(src, "bench.py", "exec"), mod.__dict__)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--model", "opt-125m", "--batch-size", "4", "--input-tokens", "80", "--max-new-tokens",
                                      "4", "--num-minibatch", "2", "--no-cpu-baseline", "--layers", "2", "--steps", "2", "--warmup", "3"])
    out = io.StringIO()
    assert mod._main(mod.parse(), out) == 0
    lines = out.getvalue().strip().splitlines()
    assert len(lines) == 1                                              # ONE JSON line
    line = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert k in line, k
    assert line["unit"] == "tokens/s" and line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] >= 3
    assert line["higher_is_better"] is True and line["scaling"] == "strong" and line["vs_baseline"] is None and line["dtype"] == "bf16"
    assert set(line["e2e"]) == {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert line["e2e"]["h2d_bytes_per_step"] == 4 * 80 * 8 and line["e2e"]["d2h_bytes_per_step"] == 4 * 84 * 8
    assert set(line["clocks"]) == {"sm_mhz", "sm_max_mhz", "reasons"}
    r = line["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and r["launches"] == 4 * 2 * 2       # 4 GEMMs x 2 layers x 2 minibatches
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert line["roofline_decode"]["bound"] == "hbm" and "workload" in line["config"] and "model" not in line["config"]
    # the parity block compares the timed model's own outputs with the oracle; on the kernel stand-in (same PyTorch ops, one
    # thread) the two are the same arithmetic, so the comparison must come out exact
    par = line["parity"]
    assert par["prefill_hidden_rel_err"] == 0.0 and par["tokens_equal_frac"] == 1.0 and par["sequences_identical"] == 4
    assert par["first_divergences"] == 0 and par["tokens_compared"] == 4 * 4


# ------------------------------------------------------------------ placement knobs end to end: streamed layers, spilled K/V

class _FakeArena:
    """Stand-in for streamer.HostArena (pinned host memory needs a CUDA driver)."""

    def __init__(self, numel):
        self.nbytes = int(numel) * 2
        self.tensor = torch.zeros(int(numel), dtype=BF16)

    def close(self):
        self.tensor = None


class _FakeStreamerLib:
    """Executes lia_streamer_prefetch as an immediate copy host slab -> slot, so a slot always holds what the schedule put
    there LAST: a layer computed from the wrong slot, or from a slot that was recycled too early, changes the tokens."""

    def __init__(self):
        self.streamers = []
        self.prefetches = 0

    def lia_streamer_create(self, arr, n, nbytes):
        return len(self.streamers) + 1

    def _st(self, h):
        return self.streamers[h - 1]

    def lia_streamer_prefetch(self, h, slot, ptr, nbytes):
        st = self._st(h)
        src = next(t for t in st.host if t.data_ptr() == ptr)
        st.slots[slot].copy_(src)
        self.prefetches += 1
        return 0

    def lia_streamer_wait(self, h, slot, stream):
        return 0

    def lia_streamer_release(self, h, slot, stream):
        return 0

    def lia_streamer_destroy(self, h):
        return 0

    def lia_streamer_stats(self, h, b, ms):
        return 0


class _EagerStream:
    def wait_event(self, ev):
        pass

    def synchronize(self):
        pass


class _NoEvent:
    def record(self, stream=None):
        pass


@pytest.fixture
def cpu_placement(cpu_ops, monkeypatch):
    """On top of the kernel stand-in: pinned arenas, the layer streamer's library calls and the spill's streams/events
    replaced by eager CPU equivalents, so generate() runs with streamed layers and spilled K/V in the CPU suite."""
    from lia_b200 import _lib, kv_spill, modeling_opt, streamer
    real_load = _lib.load
    fake = _FakeStreamerLib()

    class Lib:
        def __getattr__(self, name):
            return getattr(fake, name) if name.startswith("lia_streamer_") else getattr(real_load(), name)
    monkeypatch.setattr(_lib, "load", lambda: Lib())
    monkeypatch.setattr(streamer, "HostArena", _FakeArena)
    monkeypatch.setattr(modeling_opt, "HostArena", _FakeArena)
    monkeypatch.setattr(kv_spill, "HostArena", _FakeArena)
    real_init = streamer.LayerStreamer.__init__

    def init(self, layout, host_slabs, device):
        real_init(self, layout, host_slabs, device)
        fake.streamers.append(self)
    monkeypatch.setattr(streamer.LayerStreamer, "__init__", init)

    class S:
        cuda_stream = 0
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a: S())
    monkeypatch.setattr(kv_spill.KVSpill, "_new_stream", lambda self: _EagerStream())
    monkeypatch.setattr(kv_spill.KVSpill, "_new_event", lambda self: _NoEvent())
    monkeypatch.setattr(kv_spill.KVSpill, "_current_stream", lambda self: _EagerStream())
    monkeypatch.setattr(kv_spill.KVSpill, "_copy_async", lambda self, dst, src: dst.copy_(src))
    yield fake
    for st in fake.streamers:          # their handles are fake: they must never reach the real lia_streamer_destroy
        st.handle = None


@pytest.mark.parametrize("pct,kv_res,nmb,no_overlap", [(40, None, 2, False), (0, None, 1, False), (100, 2, 2, False), (40, 1, 2, False),
                                                       (0, 0, 1, False), (40, 1, 2, True), (0, 3, 1, True), (20, 4, 3, False)])
def test_streamed_layers_and_spilled_kv_do_not_change_results(cpu_placement, pct, kv_res, nmb, no_overlap):
    """gpu_percentage (layers streamed through two device slots), kv_resident_layers (K/V of the last layers spilled to
    host memory and passed through two device slots), num_minibatch and --no-overlap are placement / scheduling knobs:
    tokens and the cache handed back through past_key_values must equal the fully resident run's, call after call."""
    cfg = lia_b200.OPTConfig(hidden_size=128, num_hidden_layers=5, num_attention_heads=2, ffn_dim=256, vocab_size=320,
                             max_position_embeddings=48)
    B, S, new = 4, 7, 5
    ids = torch.randint(3, cfg.vocab_size, (B, S), generator=torch.Generator().manual_seed(3))
    ref_m = lia_b200.OPTForCausalLM(cfg, "cpu").init_weights(seed=9, bias_std=0.05, ln_std=0.1)
    ref_m.use_cuda_graphs = False
    ref = ref_m.generate(ids, max_new_tokens=new, min_new_tokens=new)
    ref_st = next(iter(ref_m._states.values()))
    T = S + new - 1
    m = lia_b200.OPTForCausalLM(cfg, "cpu").init_weights(seed=9, bias_std=0.05, ln_std=0.1, gpu_percentage=pct)
    m.use_cuda_graphs = False
    m.kv_resident_layers = kv_res
    dec = m.model.decoder
    assert dec.n_resident == (5 if pct >= 100 else int(5 * pct / 100)) and (dec.streamer is None) == (pct >= 100)
    for rep in range(3):                           # state re-use: wrap-around prefetches of the previous call
        tok = m.generate(ids, max_new_tokens=new, min_new_tokens=new, num_minibatch=nmb, gpu_percentage=pct, no_overlap=no_overlap)
        assert torch.equal(tok, ref), (rep, pct, kv_res)
        st = next(iter(m._states.values()))
        assert st.kv_resident == (5 if kv_res is None else kv_res) and (st.spill is None) == (kv_res is None)
        for li, p in enumerate(st.past_key_values(T)):
            assert torch.equal(p[1][:T], ref_st.kc[li][:T]) and torch.equal(p[2][:T], ref_st.vc[li][:T]), (rep, li)
    if dec.streamer is not None:
        n_str = 5 - dec.n_resident
        assert cpu_placement.prefetches >= n_str * new        # every streamed layer crosses the "link" once per forward
        if no_overlap and n_str > 1:
            assert dec.streamer.loaded[1:] == [-1] * (dec.streamer.n_slots - 1)      # only slot 0 was ever used
    with pytest.raises(ValueError, match="gpu_percentage"):
        m.generate(ids, max_new_tokens=new, gpu_percentage=(pct + 50) % 100 + 1 if pct != 100 else 50)


def test_host_layer_pool_aliases_streamed_layers(cpu_placement, monkeypatch):
    """LIA_HOST_LAYER_POOL (host RAM smaller than the streamed weights): streamed layer j aliases pinned slab j % pool."""
    monkeypatch.setenv("LIA_HOST_LAYER_POOL", "2")
    cfg = lia_b200.OPTConfig(hidden_size=128, num_hidden_layers=6, num_attention_heads=2, ffn_dim=256, vocab_size=320,
                             max_position_embeddings=48)
    m = lia_b200.OPTForCausalLM(cfg, "cpu").init_weights(seed=9, gpu_percentage=34)      # int(6 * 0.34) = 2 resident
    dec = m.model.decoder
    assert dec.n_resident == 2 and dec.host_pool == 2 and len(dec.host_slabs) == 4
    assert dec.host_slabs[0].data_ptr() == dec.host_slabs[2].data_ptr() and dec.host_slabs[1].data_ptr() == dec.host_slabs[3].data_ptr()
    assert dec.host_arena.nbytes == 2 * m.layout.nbytes
    m.use_cuda_graphs = False
    ids = torch.randint(3, cfg.vocab_size, (2, 5), generator=torch.Generator().manual_seed(3))
    assert m.generate(ids, max_new_tokens=3, gpu_percentage=34).shape == (2, 8)


@pytest.mark.parametrize("golden", ["model_hf_tiny", "model_hf_postln_tiny"])
def test_from_pretrained_forms_on_cpu(cpu_placement, golden_dir, tmp_path, golden):
    """from_pretrained (run_generation.py:159-167) over an HF safetensors directory and over the native slab directory,
    resident and fully streamed, for the plain model and for opt-350m's shape: same slabs as load_state_dict and the
    greedy tokens of the oracle on those weights (the goldens' stock-transformers tokens where bf16 does not hit a tie)."""
    import json
    from lia_b200 import checkpoint
    z = np.load(os.path.join(golden_dir, golden + ".npz"))
    bf = lambda a: torch.from_numpy(a.view(np.int16).copy()).view(BF16)  # noqa: E731
    sd = {k[3:]: bf(z[k]) for k in z.files if k.startswith("sd:")}
    h, L, H, V, P, e_dim, pre = (int(z[k]) for k in ("h", "L", "H", "V", "P", "word_dim", "pre_ln"))
    hf = tmp_path / "hf"
    hf.mkdir()
    checkpoint.write_safetensors(str(hf / "model.safetensors"), sd)
    json.dump({"model_type": "opt", "hidden_size": h, "num_hidden_layers": L, "num_attention_heads": H, "ffn_dim": 4 * h,
               "vocab_size": V, "max_position_embeddings": P, "word_embed_proj_dim": e_dim, "do_layer_norm_before": bool(pre)},
              open(hf / "config.json", "w"))
    checkpoint.convert(str(hf), str(tmp_path / "slabs"))
    ids = torch.from_numpy(z["input_ids"])
    new = int(z["new"])
    cfg = lia_b200.OPTConfig(hidden_size=h, num_hidden_layers=L, num_attention_heads=H, ffn_dim=4 * h, vocab_size=V,
                             max_position_embeddings=P, do_layer_norm_before=bool(pre), word_embed_proj_dim=0 if e_dim == h else e_dim)
    base = lia_b200.OPTForCausalLM(cfg, "cpu").load_state_dict(sd)
    base.use_cuda_graphs = False
    want = base.generate(ids, max_new_tokens=new, min_new_tokens=new)
    with torch.no_grad():
        assert torch.equal(want, opt_ref.greedy_generate(oracle_model(base), ids, new))
    if golden == "model_hf_tiny":
        assert np.array_equal(want.numpy(), z["tokens"])
    for d in ("hf", "slabs"):
        for pct in (100, 0):
            m = lia_b200.OPTForCausalLM.from_pretrained(str(tmp_path / d), "cpu", gpu_percentage=pct)
            m.use_cuda_graphs = False
            dec = m.model.decoder
            assert dec.n_resident == (L if pct == 100 else 0)
            assert m.config.do_layer_norm_before == bool(pre) and m.config.embed_dim == e_dim
            for i in range(L):
                got = dec.resident[i] if pct == 100 else dec.host_slabs[i]
                assert torch.equal(got, base.model.decoder.resident[i]), (d, pct, i)
            assert torch.equal(m.generate(ids, max_new_tokens=new, min_new_tokens=new, gpu_percentage=pct), want), (d, pct)
    with pytest.raises(ValueError, match="tensor-parallel world"):
        lia_b200.OPTForCausalLM.from_pretrained(str(tmp_path / "slabs"), "cpu", tp_rank=0, tp_world=2)


def test_operator_registry_host_logic(cpu_ops, monkeypatch):
    """The reference's op-level plugin API (ipex.llm.modules, llm/modules/utils.py:24-93) with this build's table: call
    signatures, shape handling and the (seq_info, key_cache, value_cache, beam_idx) past of IndirectAccessKVCache, on the
    kernel stand-in (tests/cpu/test_ipex_llm_module.py:166,200 and tests/cpu/test_masked_mha.py are the reference's own
    checks of these ops)."""
    from lia_b200 import llm_modules as lm
    monkeypatch.setattr(lm, "DEVICE", "cpu")
    assert set(lm.fusion_modules["cuda"]) == {lm.IPEXCustomOpType.LINEAR_RELU, lm.IPEXCustomOpType.LINEAR_ADD,
                                              lm.IPEXCustomOpType.FAST_LAYERNORM, lm.IPEXCustomOpType.INDIRECTACCESS_KVCACHE}
    assert {t.name: t.value for t in lm.IPEXCustomOpType} == {"LINEAR_RELU": 3, "LINEAR_ADD": 6, "FAST_LAYERNORM": 12,
                                                              "INDIRECTACCESS_KVCACHE": 14}          # utils.py:24-40
    torch.manual_seed(0)
    lin = torch.nn.Linear(64, 96).to(BF16)
    x, y = torch.randn(2, 3, 64).to(BF16), torch.randn(2, 3, 96).to(BF16)
    with torch.no_grad():
        want = torch.matmul(x, lin.weight.t()) + lin.bias          # two roundings, as the GPU branch's eager ops (A:393-394)
        assert torch.equal(lm.LinearRelu(lin)(x), torch.relu(want))
        assert torch.equal(lm.LinearAdd(lin)(x, y), y + want) and lm.LinearAdd(lin)(x, y).shape == (2, 3, 96)
        ln = torch.nn.LayerNorm(64).to(BF16)
        ln.weight.normal_(1, 0.1)
        ln.bias.normal_(0, 0.1)
        ref_ln = torch.nn.functional.layer_norm(x, (64,), ln.weight, ln.bias, 1e-5)
        assert torch.equal(lm.FastLayerNorm(64, 1e-5, ln.weight, ln.bias)(x), ref_ln)
        assert torch.equal(lm.FastLayerNorm.apply(x, 64, ln.weight, ln.bias, 1e-5), ref_ln)
    B, S, H, d = 2, 5, 2, 64
    q, k, v = (torch.randn(B, S, H, d).to(BF16) for _ in range(3))
    cache = lm.IndirectAccessKVCache(text_max_length=16)
    out, weights_, past = cache(q, k, v, d ** 0.5, None, None, None)

    def naive(q_, k_, v_, causal):
        s = (q_.float().permute(0, 2, 1, 3) @ k_.float().permute(0, 2, 3, 1)) / d ** 0.5
        if causal:
            s = s.masked_fill(torch.triu(torch.ones(s.shape[-2:], dtype=torch.bool), 1), float("-inf"))
        return (torch.softmax(s, -1) @ v_.float().permute(0, 2, 1, 3)).permute(0, 2, 1, 3)
    assert weights_ is None and out.shape == (B, S, H, d) and past[0].shape[2] == S and past[1].shape == (16, B, H, d)
    assert (out.float() - naive(q, k, v, True)).abs().max() <= 2e-2 * naive(q, k, v, True).abs().max()
    ks, vs = k, v
    for step in range(2):
        q1, k1, v1 = (torch.randn(B, 1, H, d).to(BF16) for _ in range(3))
        out, _, past = cache(q1, k1, v1, d ** 0.5, past, None, None)
        ks, vs = torch.cat([ks, k1], 1), torch.cat([vs, v1], 1)
        ref = naive(q1, ks, vs, False)
        assert past[0].shape[2] == S + step + 1 and (out.float() - ref).abs().max() <= 2e-2 * ref.abs().max()
    assert torch.equal(past[1][:S + 2].permute(1, 0, 2, 3), ks) and past[3].shape == (16, B)
    with pytest.raises(NotImplementedError):
        cache(q, k, v, d ** 0.5, None, torch.ones(1), None)                 # head_mask
