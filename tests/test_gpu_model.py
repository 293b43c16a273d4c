"""Path-level parity on a real B200, through the reference-shaped Python faces (which call the
C ABI): golden vectors made from the reference's own layer code, the oracle on the same
seeded inputs, stock-transformers tokens, and size-independent properties at full size.

North-star bars (BASELINE.json): per-layer hidden-state max relative error <= 1e-2,
identical greedy tokens for the first 32 generated."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16
REL_TOL = 1e-2     # north_star: per-layer hidden-state max relative error


@pytest.fixture(scope="module")
def lia():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import lia_b200
    return lia_b200


def _bf16(a):
    return torch.from_numpy(a.view(np.int16).copy()).view(BF16)


def rel_err(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


def _single_layer_model(lia, h, H, f, w, V=64, P=64):
    from lia_b200.weights import random_embeddings
    cfg = lia.OPTConfig(hidden_size=h, num_hidden_layers=1, num_attention_heads=H, ffn_dim=f, vocab_size=V,
                        max_position_embeddings=P)
    m = lia.OPTForCausalLM(cfg, "cuda")
    m.model.decoder.load_embeddings(random_embeddings(V, h, P, 1, "cuda"))
    m.model.decoder.load_layers(lambda i, dev: {k: t.to(dev) for k, t in w.items()}, 100)
    return m


@pytest.mark.parametrize("name", ["layer_d64", "layer_d128", "layer_ragged"])
def test_layer_vs_reference_golden(lia, golden_dir, name):
    """decoder_layer(...) face against outputs of the reference's own OPTDecoderLayer_forward."""
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    B, S, h, H, new = (int(z[k]) for k in ("B", "S", "h", "H", "new"))
    w = {k[2:]: _bf16(z[k]) for k in z.files if k.startswith("w_")}
    m = _single_layer_model(lia, h, H, 4 * h, w)
    layer = m.model.decoder.layers[0]
    past = None
    for step in range(new + 1):
        x = _bf16(z[f"x{step}"]).cuda()
        out = layer(x, attention_mask=None, past_key_value=past, use_cache=True, policy=3, max_new_tokens=new)
        y, past = out[0], out[1]
        assert past[0].shape[2] == S + step
        e = rel_err(y, _bf16(z[f"y{step}"]))
        assert e <= REL_TOL, (name, step, e)
    T = S + new
    assert rel_err(past[1][:T], _bf16(z["kcache"])) <= REL_TOL
    assert rel_err(past[2][:T], _bf16(z["vcache"])) <= REL_TOL


def test_layer_policy0_returns_new_kv_rows(lia, golden_dir):
    z = np.load(os.path.join(golden_dir, "layer_d64.npz"))
    B, S, h, H, new = (int(z[k]) for k in ("B", "S", "h", "H", "new"))
    w = {k[2:]: _bf16(z[k]) for k in z.files if k.startswith("w_")}
    m = _single_layer_model(lia, h, H, 4 * h, w)
    out = m.model.decoder.layers[0](_bf16(z["x0"]).cuda(), use_cache=True, policy=0, max_new_tokens=new)
    assert len(out) == 4 and out[2].shape == (S, B, H, h // H)          # decoder.py:331-333
    # gpu_layer (16-entry list, lia/modeling_opt.py:272-293) gives the same result as resident weights
    from lia_b200.weights import LAYER_KEYS
    gl = [w[k].cuda() for k in LAYER_KEYS]
    out2 = m.model.decoder.layers[0](_bf16(z["x0"]).cuda(), use_cache=True, policy=0, max_new_tokens=new, gpu_layer=gl)
    assert torch.equal(out[0], out2[0])


def test_tiny_model_tokens_match_stock_transformers(lia, golden_dir):
    z = np.load(os.path.join(golden_dir, "model_hf_tiny.npz"))
    sd = {k[3:]: _bf16(z[k]) for k in z.files if k.startswith("sd:")}
    cfg = lia.OPTConfig(hidden_size=int(z["h"]), num_hidden_layers=int(z["L"]), num_attention_heads=int(z["H"]),
                        ffn_dim=4 * int(z["h"]), vocab_size=int(z["V"]), max_position_embeddings=int(z["P"]))
    m = lia.OPTForCausalLM(cfg, "cuda").load_state_dict(sd)
    ids = torch.from_numpy(z["input_ids"])
    new = int(z["new"])
    # forward face: logits within the reference's own nightly tolerance (0.1) of HF fp32
    logits, past = m(input_ids=ids.cuda(), attention_mask=torch.ones_like(ids), max_new_tokens=new, prefill_policy=0,
                     decoding_policy=0, gpu_percentage=100, num_minibatch=1)
    assert logits.shape == (ids.shape[0], 1, int(z["V"])) and past[0][0].shape[2] == ids.shape[1]
    assert np.abs(logits[:, 0].float().cpu().numpy() - z["prefill_last_logits"]).max() < 0.1
    for rep in range(3):          # eager, graph capture, graph replay
        toks = m.generate(ids, max_new_tokens=new, min_new_tokens=new, do_sample=False, num_beams=1)
        assert np.array_equal(toks.numpy(), z["tokens"]), rep


def test_from_pretrained_checkpoint_forms(lia, golden_dir, tmp_path):
    """from_pretrained (run_generation.py:159-167) over an HF safetensors directory and over this build's native
    slab directory, resident and streamed: same bits in HBM as load_state_dict, same tokens as stock transformers."""
    from lia_b200 import checkpoint
    z = np.load(os.path.join(golden_dir, "model_hf_tiny.npz"))
    sd = {k[3:]: _bf16(z[k]) for k in z.files if k.startswith("sd:")}
    h, L, H, V, P = (int(z[k]) for k in ("h", "L", "H", "V", "P"))
    hf = tmp_path / "hf"
    hf.mkdir()
    checkpoint.write_safetensors(str(hf / "model.safetensors"), sd)
    import json
    json.dump({"model_type": "opt", "hidden_size": h, "num_hidden_layers": L, "num_attention_heads": H, "ffn_dim": 4 * h,
               "vocab_size": V, "max_position_embeddings": P, "word_embed_proj_dim": h}, open(hf / "config.json", "w"))
    checkpoint.convert(str(hf), str(tmp_path / "slabs"))
    ids = torch.from_numpy(z["input_ids"])
    new = int(z["new"])
    base = lia.OPTForCausalLM(lia.OPTConfig(hidden_size=h, num_hidden_layers=L, num_attention_heads=H, ffn_dim=4 * h, vocab_size=V,
                                            max_position_embeddings=P), "cuda").load_state_dict(sd)
    for d in ("hf", "slabs"):
        for pct in (100, 0):
            m = lia.OPTForCausalLM.from_pretrained(str(tmp_path / d), "cuda", gpu_percentage=pct)
            dec = m.model.decoder
            assert dec.n_resident == (L if pct == 100 else 0)
            for i in range(L):
                got = dec.resident[i] if pct == 100 else dec.host_slabs[i]
                assert torch.equal(got.cpu(), base.model.decoder.resident[i].cpu()), (d, pct, i)
            toks = m.generate(ids, max_new_tokens=new, min_new_tokens=new, gpu_percentage=pct)
            assert np.array_equal(toks.numpy(), z["tokens"]), (d, pct)
            del m
    with pytest.raises(ValueError, match="tensor-parallel world"):
        lia.OPTForCausalLM.from_pretrained(str(tmp_path / "slabs"), "cuda", tp_rank=0, tp_world=2)


def _oracle_model(m, device):
    """Unpack a lia_b200 model's slabs into the oracle's dict format (same bits)."""
    dec = m.model.decoder
    hq = dec.layout.hq
    layers = []
    for v in dec.resident_views:
        w = {k: v[k] for k in ("ln1_w", "ln1_b", "o_w", "o_b", "ln2_w", "ln2_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b")}
        w["q_w"], w["k_w"], w["v_w"] = v["qkv_w"][:hq], v["qkv_w"][hq:2 * hq], v["qkv_w"][2 * hq:]
        w["q_b"], w["k_b"], w["v_b"] = v["qkv_b"][:hq], v["qkv_b"][hq:2 * hq], v["qkv_b"][2 * hq:]
        layers.append({k: t.to(device) for k, t in w.items()})
    return {"H": m.config.num_attention_heads, "layers": layers, "embed_tokens": dec.embed_tokens.to(device),
            "embed_positions": dec.embed_positions.to(device), "final_ln_w": dec.final_ln_w.to(device),
            "final_ln_b": dec.final_ln_b.to(device)}


def _ulp(x):
    return 2.0 ** (torch.floor(torch.log2(x.abs().clamp_min(1e-30))) - 7)


@pytest.mark.parametrize("cfgname,L,B,S,new,nmb", [("opt-1.3b", 3, 8, 256, 32, 1), ("opt-30b", 2, 4, 64, 32, 2)])
def test_greedy_tokens_and_hidden_vs_oracle(lia, cfgname, L, B, S, new, nmb):
    """Real layer dims, reduced depth, against the oracle run on the same GPU (= the reference's eager
    op sequence on cuBLAS/ATen).

    * per-layer hidden state (each layer fed the oracle's input): max rel err <= 1e-2
    * 32 greedy tokens: identical, except that a sequence may leave the oracle's path at a step where
      the oracle's own top-2 logits are within bf16 resolution of each other (random-init logits are
      nearly flat, SURVEY.md section 7 "hard parts": fp32 summation order alone flips such near-ties;
      the oracle itself flips them between CPU and GPU).  Every divergence is checked to be such a
      near-tie; anything else fails."""
    from oracle import opt_ref
    cfg = lia.modeling_opt.get_config(cfgname)
    cfg.num_hidden_layers = L
    m = lia.OPTForCausalLM(cfg, "cuda").init_weights(seed=3, bias_std=0.02, ln_std=0.05)
    om = _oracle_model(m, "cuda")
    g = torch.Generator().manual_seed(1234)
    ids = torch.randint(3, cfg.vocab_size, (B, S), generator=g)
    ones = torch.ones(B, S, dtype=torch.long, device="cuda")
    with torch.no_grad():
        ref_logits = []
        ref = opt_ref.greedy_generate(om, ids.cuda(), new, collect_logits=ref_logits)
        hs = []
        cache = opt_ref.new_cache(om, B, S + new)
        opt_ref.decoder_forward(om, ids.cuda(), ones, cache, 0, collect=hs)
        x0 = opt_ref.embed(om, ids.cuda(), ones, 0)
    chained = x0
    for li, layer in enumerate(m.model.decoder.layers):
        inp = x0 if li == 0 else hs[li - 1]
        y = layer(inp, use_cache=True, policy=3, max_new_tokens=new)[0]
        assert rel_err(y, hs[li]) <= REL_TOL, (li, rel_err(y, hs[li]))
        chained = layer(chained, use_cache=True, policy=3, max_new_tokens=new)[0]
    assert rel_err(chained, hs[-1]) <= 3 * REL_TOL          # errors may add up across layers, not blow up
    for rep in range(3):                                    # eager, graph capture, graph replay
        toks = m.generate(ids, max_new_tokens=new, min_new_tokens=new, num_minibatch=nmb, prefill_policy=0,
                          decoding_policy=0).cpu()
        if rep == 0:
            first = toks
        assert torch.equal(toks, first), "generate() is not deterministic across eager / graph replays"
    assert torch.equal(toks[:, :S], ids)
    n_ident = 0
    for b in range(B):
        diff = torch.nonzero(toks[b] != ref[b].cpu()).flatten()
        if diff.numel() == 0:
            n_ident += 1
            continue
        t = int(diff[0]) - S                                # first generated token that differs
        lg = ref_logits[t][b].float().cpu()
        lg[cfg.eos_token_id] = float("-inf")
        margin = (lg[ref[b, S + t]] - lg[toks[b, S + t]]).item()
        tol = 3 * _ulp(lg[ref[b, S + t]]).item()
        assert 0 <= margin <= tol, (f"sequence {b} leaves the oracle at generated token {t} where the oracle's margin "
                                    f"{margin:.4f} is not a bf16 near-tie (tol {tol:.4f})")
    print(f"{cfgname}: {n_ident}/{B} sequences identical for all {new} tokens; the rest diverge at a bf16 near-tie")


@pytest.mark.parametrize("cfgname,L,B,S,new", [("opt-1.3b", 3, 8, 256, 32), ("opt-30b", 2, 8, 64, 32)])
def test_teacher_forced_tokens_vs_oracle(lia, cfgname, L, B, S, new):
    """EVERY one of the B x new greedy decisions, not only those up to a sequence's first divergence: the oracle generates,
    and this build is driven through its reference-shaped forward face (models.py:371-445) with the ORACLE's token at every
    step, so both always see the same context.  Bar: identical argmax wherever the oracle's own top-2 margin exceeds 3 bf16
    ulps of the winning logit; below that, fp32 summation order decides (the oracle flips such ties between CPU and GPU).
    Measured on hardware (profiles/r2/r2_token_parity_*.json): 246/256 and 254/256 identical at these two shapes with every
    flip at a margin of at most 1 ulp; at the full OPT-30B bench config (48 layers, 64 x 32 decisions) 1962/2048 identical
    with the worst flip at 4 ulps -- the rounding noise of 48 chained bf16 layers, see the per-layer errors in bench.py's
    parity block."""
    from oracle import opt_ref
    cfg = lia.modeling_opt.get_config(cfgname)
    cfg.num_hidden_layers = L
    m = lia.OPTForCausalLM(cfg, "cuda").init_weights(seed=3, bias_std=0.02, ln_std=0.05)
    om = _oracle_model(m, "cuda")
    ids = torch.randint(3, cfg.vocab_size, (B, S), generator=torch.Generator().manual_seed(1234)).cuda()
    ref_logits = []
    with torch.no_grad():
        ref = opt_ref.greedy_generate(om, ids, new, collect_logits=ref_logits)
    mask = torch.ones(B, S, dtype=torch.long, device="cuda")
    logits, past = m(input_ids=ids, attention_mask=mask, max_new_tokens=new, prefill_policy=0, decoding_policy=0)
    flips, worst = 0, 0.0
    for t in range(new):
        lg = logits[:, -1].float()
        lg[:, cfg.eos_token_id] = float("-inf")
        ours = lg.argmax(-1)
        want = ref[:, S + t]
        rl = ref_logits[t].float().clone()
        rl[:, cfg.eos_token_id] = float("-inf")
        for b in (ours != want).nonzero().flatten().tolist():
            margin = (rl[b, want[b]] - rl[b, ours[b]]).item()
            ulps = margin / _ulp(rl[b, want[b]]).item()
            flips += 1
            worst = max(worst, ulps)
            assert 0 <= ulps <= 3.0, (f"decision (sequence {b}, token {t}) differs from the oracle's although its top-2 margin is "
                                      f"{ulps:.2f} bf16 ulps")
        if t + 1 < new:
            mask = torch.cat([mask, mask.new_ones(B, 1)], dim=-1)
            logits, past = m(input_ids=want[:, None].contiguous(), attention_mask=mask, past_key_values=past, max_new_tokens=new)
    assert flips <= 0.1 * B * new, f"{flips} of {B * new} decisions differ"
    print(f"{cfgname}: {B * new - flips}/{B * new} teacher-forced decisions identical, worst flip margin {worst:.2f} ulp")


def test_padded_prompts_follow_the_reference_mask_semantics(lia):
    """Ragged (padded) prompts.  On the reference's GPU branch the attention mask moves only the learned positions
    (cumsum rule, lia/modeling_opt.py:368-378): attention is pure-causal in prefill and unmasked in decode
    (attentions.py:446-449, 500), and generate() derives the mask from pad ids when none is passed
    (lia/generation_utils.py:469-485).  Hidden states vs the oracle fed the same mask; generate() with the implicit
    and the explicit mask must agree bit for bit and react to the mask."""
    from oracle import opt_ref
    cfg = lia.OPTConfig(hidden_size=256, num_hidden_layers=2, num_attention_heads=2, ffn_dim=1024, vocab_size=512,
                        max_position_embeddings=64)
    m = lia.OPTForCausalLM(cfg, "cuda").init_weights(seed=4, bias_std=0.02, ln_std=0.05)
    om = _oracle_model(m, "cuda")
    B, S, new = 5, 12, 4
    ids = torch.randint(3, cfg.vocab_size, (B, S), generator=torch.Generator().manual_seed(5))
    ids[0, :4] = cfg.pad_token_id          # left padding
    ids[1, :1] = cfg.pad_token_id
    ids[2, -3:] = cfg.pad_token_id         # right padding
    ids[3, 6] = cfg.pad_token_id           # a hole
    mask = ids.ne(cfg.pad_token_id).long()
    dec = m.model.decoder
    hidden, past = dec(input_ids=ids.cuda(), attention_mask=mask.cuda(), max_new_tokens=new)
    nxt = torch.randint(3, cfg.vocab_size, (B, 1), generator=torch.Generator().manual_seed(6))
    mask1 = torch.cat([mask, mask.new_ones(B, 1)], 1)
    hidden1, _ = dec(input_ids=nxt.cuda(), attention_mask=mask1.cuda(), past_key_values=past, max_new_tokens=new)
    with torch.no_grad():
        cache = opt_ref.new_cache(om, B, S + new)
        href = opt_ref.decoder_forward(om, ids.cuda(), mask.cuda(), cache, 0)
        href1 = opt_ref.decoder_forward(om, nxt.cuda(), mask1.cuda(), cache, S)
        hones = opt_ref.decoder_forward(om, ids.cuda(), torch.ones_like(mask).cuda(), opt_ref.new_cache(om, B, S + new), 0)
    assert rel_err(hidden, href) <= 3 * REL_TOL and rel_err(hidden1, href1) <= 3 * REL_TOL      # chained through 2 layers
    assert rel_err(hidden, hones) > 10 * REL_TOL                      # the mask matters: all-ones positions are far off
    with pytest.raises(ValueError):
        dec(input_ids=ids.cuda(), attention_mask=mask1.cuda(), max_new_tokens=new)               # M:1127-1131
    kw = dict(max_new_tokens=new, min_new_tokens=new, prefill_policy=0, decoding_policy=0)
    outs = [m.generate(ids, **kw) for _ in range(3)]                  # implicit mask; eager, capture, replay
    outs += [m.generate(ids, attention_mask=mask, **kw)]
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    with torch.no_grad():
        ref_logits = []
        ref = opt_ref.greedy_generate(om, ids.cuda(), new, collect_logits=ref_logits).cpu()
    for b in range(B):
        diff = torch.nonzero(outs[0][b] != ref[b]).flatten()
        if diff.numel():                                              # only a bf16 near-tie may flip a token
            t = int(diff[0]) - S
            lg = ref_logits[t][b].float().cpu()
            lg[cfg.eos_token_id] = float("-inf")
            margin = (lg[ref[b, S + t]] - lg[outs[0][b, S + t]]).item()
            assert 0 <= margin <= 3 * _ulp(lg[ref[b, S + t]]).item(), (b, t, margin)
    # after an unpadded call on the same shape the mask state must not leak
    clean = ids.clamp(min=3)
    a = m.generate(clean, **kw)
    assert torch.equal(a, m.generate(clean, attention_mask=torch.ones_like(mask), **kw))


def test_scheduling_knobs_do_not_change_results(lia):
    """gpu_percentage (streaming) and num_minibatch are scheduling knobs: outputs must be bit-identical
    (the reference's minibatch quirk, SURVEY.md A.4, is not reproduced)."""
    cfg = lia.modeling_opt.get_config("opt-1.3b")
    cfg.num_hidden_layers = 5
    g = torch.Generator().manual_seed(7)
    ids = torch.randint(3, cfg.vocab_size, (8, 64), generator=g)
    outs = []
    for pct, nmb in [(100, 1), (100, 2), (40, 2), (0, 1), (20, 1)]:      # all keep M > 128 rows per GEMM in prefill
        m = lia.OPTForCausalLM(cfg, "cuda").init_weights(seed=5, gpu_percentage=pct)
        assert m.model.decoder.n_resident == (5 if pct == 100 else int(5 * pct / 100))
        for rep in range(3):
            outs.append(m.generate(ids, max_new_tokens=8, min_new_tokens=8, num_minibatch=nmb, gpu_percentage=pct,
                                   pin_weight=True, prefill_policy=0, decoding_policy=0))
        if pct < 100:
            st = m.model.decoder.streamer.stats()
            assert st["bytes"] >= 3 * (5 - m.model.decoder.n_resident) * m.layout.nbytes
        del m
        torch.cuda.empty_cache()
    for i, o in enumerate(outs[1:]):
        assert torch.equal(o, outs[0]), i + 1


def test_kv_spill_does_not_change_results(lia):
    """KV-cache host spill (kv_spill.py; reference load_kv_cache / store_cache / store_cache_decoding,
    lia/modeling_opt.py:326-349) is a placement knob: tokens and the cache contents handed back through
    past_key_values must be bit-identical whichever layers' K/V live in pinned host memory, also together
    with weight streaming."""
    cfg = lia.modeling_opt.get_config("opt-1.3b")
    cfg.num_hidden_layers = 5
    g = torch.Generator().manual_seed(11)
    B, S, new = 8, 64, 8
    ids = torch.randint(3, cfg.vocab_size, (B, S), generator=g)
    ref_tok = ref_kv = None
    for pct, kv_res, nmb in [(100, None, 1), (100, 0, 1), (100, 2, 2), (100, 4, 1), (40, 1, 2), (0, 3, 1)]:
        m = lia.OPTForCausalLM(cfg, "cuda").init_weights(seed=5, gpu_percentage=pct)
        m.kv_resident_layers = kv_res
        for rep in range(3):                       # state re-use across generate() calls (wrap-around prefetch)
            tok = m.generate(ids, max_new_tokens=new, min_new_tokens=new, num_minibatch=nmb, gpu_percentage=pct,
                             prefill_policy=0, decoding_policy=0)
            st = next(iter(m._states.values()))
            assert st.kv_resident == (5 if kv_res is None else kv_res) and (st.spill is None) == (kv_res is None)
            T = S + new - 1                        # rows written: the last token's K/V is never needed
            kv = [(p[1][:T].cpu().clone(), p[2][:T].cpu().clone()) for p in st.past_key_values(T)]
            if ref_tok is None:
                ref_tok, ref_kv = tok, kv
                assert all(k.abs().sum() > 0 for k, _ in kv)
            assert torch.equal(tok, ref_tok), (pct, kv_res, rep)
            for li, ((k, v), (rk, rv)) in enumerate(zip(kv, ref_kv)):
                assert torch.equal(k, rk) and torch.equal(v, rv), (pct, kv_res, rep, li)
        if st.spill is not None:
            s_ = st.spill.stats()
            n_sp = 5 - kv_res
            assert st.past_key_values(T)[4][1].device.type == "cpu" and st.past_key_values(T)[4][1].is_pinned()
            assert s_["layers"] == n_sp and s_["d2h_bytes"] == 3 * 2 * n_sp * (S + new - 1) * st.spill.row * 2
        del m, st
        torch.cuda.empty_cache()


def test_host_layer_pool_streams_full_byte_count(lia, monkeypatch):
    """LIA_HOST_LAYER_POOL (host RAM smaller than the streamed weights, OPT-175B): streamed layers alias `pool`
    distinct pinned slabs; the bytes crossing PCIe per forward are those of the full model and the result is that
    of the model whose streamed layer j holds the weights of streamed layer j % pool."""
    cfg = lia.modeling_opt.get_config("opt-1.3b")
    cfg.num_hidden_layers = 5
    ids = torch.randint(3, cfg.vocab_size, (4, 48), generator=torch.Generator().manual_seed(3))
    monkeypatch.setenv("LIA_HOST_LAYER_POOL", "2")
    m = lia.OPTForCausalLM(cfg, "cuda").init_weights(seed=9, gpu_percentage=20)          # 1 resident, 4 streamed
    dec = m.model.decoder
    assert dec.host_pool == 2 and len(dec.host_slabs) == 4 and dec.host_arena.nbytes == 2 * m.layout.nbytes
    assert [t.data_ptr() for t in dec.host_slabs[2:]] == [t.data_ptr() for t in dec.host_slabs[:2]]
    tok = m.generate(ids, max_new_tokens=4, min_new_tokens=4, gpu_percentage=20)
    assert dec.streamer.stats()["bytes"] >= 4 * 4 * m.layout.nbytes                       # 4 forwards x 4 streamed layers
    monkeypatch.delenv("LIA_HOST_LAYER_POOL")
    from lia_b200.weights import random_layer
    ref = lia.OPTForCausalLM(cfg, "cuda")
    ref.model.decoder.load_embeddings({"embed_tokens": dec.embed_tokens, "embed_positions": dec.embed_positions,
                                       "final_ln_w": dec.final_ln_w, "final_ln_b": dec.final_ln_b})
    src = [0, 1, 2, 1, 2]                                                                  # layer i holds generator layer src[i]
    ref.model.decoder.load_layers(lambda i, dev: random_layer(cfg.hidden_size, cfg.ffn_dim, 9 * 100003 + 1000 + src[i], dev,
                                                              "normal", cfg.init_std), 100)
    assert torch.equal(ref.generate(ids, max_new_tokens=4, min_new_tokens=4), tok)


def test_full_size_layer_properties(lia):
    """OPT-30B layer dims at the bench batch (B=64): decode output vs oracle, KV round trip, and
    batch-permutation equivariance (a size-independent property of the path)."""
    from oracle import opt_ref
    cfg = lia.modeling_opt.get_config("opt-30b")
    cfg.num_hidden_layers = 1
    m = lia.OPTForCausalLM(cfg, "cuda").init_weights(seed=9, bias_std=0.02, ln_std=0.05)
    om = _oracle_model(m, "cuda")
    B, S, new = 64, 32, 4
    g = torch.Generator().manual_seed(2)
    x = (torch.randn(B, S, cfg.hidden_size, generator=g) * 0.5).to(BF16).cuda()
    layer = m.model.decoder.layers[0]
    y, past = layer(x, use_cache=True, policy=3, max_new_tokens=new)[:2]
    kc, vc = opt_ref.new_cache(om, B, S + new)[0]
    with torch.no_grad():
        y_ref = opt_ref.layer_forward(x, om["layers"][0], cfg.num_attention_heads, kc, vc, 0)
    assert rel_err(y, y_ref) <= REL_TOL
    assert rel_err(past[1][:S], kc[:S]) <= REL_TOL and rel_err(past[2][:S], vc[:S]) <= REL_TOL
    xd = (torch.randn(B, 1, cfg.hidden_size, generator=g) * 0.5).to(BF16).cuda()
    yd, past2 = layer(xd, past_key_value=past, use_cache=True, policy=3, max_new_tokens=new)[:2]
    with torch.no_grad():
        yd_ref = opt_ref.layer_forward(xd, om["layers"][0], cfg.num_attention_heads, kc, vc, S)
    assert past2[0].shape[2] == S + 1 and rel_err(yd, yd_ref) <= REL_TOL
    # permuting the batch permutes the output rows, bit for bit
    perm = torch.randperm(B, generator=g).cuda()
    yp = layer(x[perm].contiguous(), use_cache=True, policy=3, max_new_tokens=new)[0]
    assert torch.equal(yp, y[perm])


def test_policy1_is_refused(lia):
    cfg = lia.OPTConfig(hidden_size=128, num_hidden_layers=1, num_attention_heads=2, ffn_dim=512, vocab_size=64,
                        max_position_embeddings=32)
    m = lia.OPTForCausalLM(cfg, "cuda").init_weights()
    with pytest.raises(NotImplementedError):
        m.generate(torch.randint(3, 64, (1, 4)), max_new_tokens=2, prefill_policy=1, decoding_policy=1)


def test_operator_registry_cuda_table(lia):
    """The reference's op-level plugin API (ipex.llm.modules) with a "cuda" table: same call
    signatures, each backed by one C-ABI entry point."""
    from lia_b200 import llm_modules as lm
    assert set(lm.fusion_modules["cuda"]) == {lm.IPEXCustomOpType.LINEAR_RELU, lm.IPEXCustomOpType.LINEAR_ADD,
                                              lm.IPEXCustomOpType.FAST_LAYERNORM, lm.IPEXCustomOpType.INDIRECTACCESS_KVCACHE}
    torch.manual_seed(0)
    lin = torch.nn.Linear(256, 512).to(BF16)
    x = torch.randn(1, 4, 256).to(BF16).cuda()
    y = torch.randn(1, 4, 512).to(BF16).cuda()
    ref = torch.nn.functional.linear(x.float(), lin.weight.float().cuda(), lin.bias.float().cuda())
    # tests/cpu/test_ipex_llm_module.py:166,200 compare against eager relu(linear(x)) / linear(x)+y
    assert rel_err(lm.LinearRelu(lin)(x), torch.relu(ref)) <= REL_TOL
    assert rel_err(lm.LinearAdd(lin)(x, y), ref + y.float()) <= REL_TOL
    ln = torch.nn.LayerNorm(256).to(BF16)
    with torch.no_grad():
        ln.weight.normal_(1, 0.1); ln.bias.normal_(0, 0.1)
    ref_ln = torch.nn.functional.layer_norm(x.float(), (256,), ln.weight.float().cuda(), ln.bias.float().cuda(), 1e-5)
    assert rel_err(lm.FastLayerNorm(256, 1e-5, ln.weight, ln.bias)(x), ref_ln) <= REL_TOL
    assert rel_err(lm.FastLayerNorm.apply(x, 256, ln.weight.cuda(), ln.bias.cuda(), 1e-5), ref_ln) <= REL_TOL
    # IndirectAccessKVCache: first token then two next tokens vs naive cat-KV attention (tests/cpu/test_masked_mha.py)
    B, S, H, d = 2, 9, 4, 64
    q, k, v = (torch.randn(B, S, H, d).to(BF16).cuda() for _ in range(3))
    cache = lm.IndirectAccessKVCache(text_max_length=32)
    out, _, past = cache(q, k, v, d ** 0.5, None, None, None)
    def naive(q_, k_, v_, causal):
        s = (q_.float().permute(0, 2, 1, 3) @ k_.float().permute(0, 2, 3, 1)) / d ** 0.5
        if causal:
            s = s.masked_fill(torch.triu(torch.ones(s.shape[-2:], dtype=torch.bool, device="cuda"), 1), float("-inf"))
        return (torch.softmax(s, -1) @ v_.float().permute(0, 2, 1, 3)).permute(0, 2, 1, 3)
    assert past[0].shape[2] == S and rel_err(out, naive(q, k, v, True)) <= 2e-2      # bf16 prec of test_masked_mha.py:277-280
    ks, vs = k, v
    for step in range(2):
        q1, k1, v1 = (torch.randn(B, 1, H, d).to(BF16).cuda() for _ in range(3))
        out, _, past = cache(q1, k1, v1, d ** 0.5, past, None, None)
        ks, vs = torch.cat([ks, k1], 1), torch.cat([vs, v1], 1)
        assert past[0].shape[2] == S + step + 1 and rel_err(out, naive(q1, ks, vs, False)) <= 2e-2
    assert torch.equal(past[1][:S + 2].permute(1, 0, 2, 3), ks)                      # cache contents, time-major
