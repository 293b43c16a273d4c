"""Pin the oracle: bit-exact against outputs of the reference's own layer code
(tests/golden/layer_*.npz, made by oracle/gen_golden.py from /root/reference) and
within the reference's nightly tolerance against stock transformers."""
import os

import numpy as np
import pytest
import torch

from oracle import opt_ref

LAYER_CASES = ["layer_d64", "layer_d128", "layer_ragged", "layer_postln"]


def _bf16(a):
    return torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16)


def load_layer_case(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    w = {k[2:]: _bf16(z[k]) for k in z.files if k.startswith("w_")}
    meta = {k: int(z[k]) for k in ("B", "S", "h", "H", "new")}
    meta["pre_ln"] = bool(int(z["pre_ln"])) if "pre_ln" in z.files else True     # False: opt-350m's LayerNorm placement
    xs = [_bf16(z[f"x{i}"]) for i in range(meta["new"] + 1)]
    ys = [_bf16(z[f"y{i}"]) for i in range(meta["new"] + 1)]
    return meta, w, xs, ys, _bf16(z["kcache"]), _bf16(z["vcache"])


@pytest.mark.parametrize("name", LAYER_CASES)
def test_layer_bit_exact_vs_reference(golden_dir, name):
    torch.set_num_threads(1)
    meta, w, xs, ys, kc_ref, vc_ref = load_layer_case(golden_dir, name)
    B, S, h, H, new = (meta[k] for k in ("B", "S", "h", "H", "new"))
    kc = torch.zeros(S + new, B, H, h // H, dtype=torch.bfloat16)
    vc = torch.zeros_like(kc)
    cur = 0
    for x, y_ref in zip(xs, ys):
        y = opt_ref.layer_forward(x, w, H, kc, vc, cur, meta["pre_ln"])
        cur += x.shape[1]
        assert torch.equal(y.view(torch.int16), y_ref.view(torch.int16)), name
    assert torch.equal(kc.view(torch.int16), kc_ref.view(torch.int16))
    assert torch.equal(vc.view(torch.int16), vc_ref.view(torch.int16))


HF_CASES = ["model_hf_tiny", "model_hf_postln_tiny"]    # the second: opt-350m's shape (post-LN, project_in/out, no final LN)


def load_hf_case(golden_dir, name="model_hf_tiny"):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    sd = {k[3:]: _bf16(z[k]) for k in z.files if k.startswith("sd:")}
    model = opt_ref.model_from_hf_state_dict(sd, int(z["H"]))
    return z, model


@pytest.mark.parametrize("name", HF_CASES)
def test_model_fp32_matches_stock_transformers(golden_dir, name):
    z, model = load_hf_case(golden_dir, name)
    assert model["pre_ln"] == bool(int(z["pre_ln"])) and (model["project_in"] is not None) == (int(z["word_dim"]) != int(z["h"]))
    m32 = opt_ref.model_to(model, dtype=torch.float32)
    ids = torch.from_numpy(z["input_ids"])
    logits = []
    toks = opt_ref.greedy_generate(m32, ids, int(z["new"]), collect_logits=logits)
    assert np.array_equal(toks.numpy(), z["tokens"])
    np.testing.assert_allclose(logits[0].numpy(), z["prefill_last_logits"], atol=2e-5)


@pytest.mark.parametrize("name", HF_CASES)
def test_model_bf16_within_reference_nightly_tolerance(golden_dir, name):
    # tests/cpu/test_ipex_optimize_transformers_nightly.py:237 uses prec=0.1 for bf16 vs fp32 logits
    z, model = load_hf_case(golden_dir, name)
    ids = torch.from_numpy(z["input_ids"])
    logits = []
    opt_ref.greedy_generate(model, ids, 1, collect_logits=logits)
    assert np.abs(logits[0].numpy() - z["prefill_last_logits"]).max() < 0.1


def test_tp_sharding_restatement_sums_to_full(golden_dir):
    meta, w, xs, ys, _, _ = load_layer_case(golden_dir, "layer_d64")
    x = xs[0].float()
    wf = {k: v.float() for k, v in w.items()}
    full = torch.relu(x @ wf["fc1_w"].t() + wf["fc1_b"]) @ wf["fc2_w"].t()
    parts = 0
    for r in range(2):
        s = opt_ref.shard_layer(wf, meta["H"], r, 2)
        parts = parts + torch.relu(x @ s["fc1_w"].t() + s["fc1_b"]) @ s["fc2_w"].t()
    torch.testing.assert_close(parts, full, atol=1e-4, rtol=1e-4)


def test_positions_and_mask_rule_bit_exact_vs_reference(golden_dir):
    """Padded prompts (left, right, a hole, a fully padded row): the learned-position indices and embedding rows
    of the reference's own OPTLearnedPositionalEmbedding.forward, and its mask-from-pad-ids rule."""
    z = np.load(os.path.join(golden_dir, "positions_padded.npz"))
    ids, table = torch.from_numpy(z["ids"]), _bf16(z["table"])
    mask = opt_ref.prepare_attention_mask(ids, 1, 2)
    assert np.array_equal(mask.numpy(), z["mask"])
    assert np.array_equal(opt_ref.prepare_attention_mask(ids.clamp(min=3), 1, 2).numpy(), z["mask_nopad"])
    assert np.array_equal(opt_ref.prepare_attention_mask(ids, 1, 1).numpy(), z["mask_pad_is_eos"])
    B, S, new = int(z["B"]), int(z["S"]), int(z["new"])
    model = {"embed_tokens": torch.zeros(4, table.shape[1], dtype=torch.bfloat16), "embed_positions": table}
    full = mask
    for step in range(new + 1):
        past = 0 if step == 0 else S + step - 1
        pos = opt_ref.positions_from_mask(full, past)
        assert np.array_equal(pos.numpy(), z[f"pos{step}"]), step
        rows = opt_ref.embed(model, torch.zeros(B, pos.shape[1], dtype=torch.long), full, past)   # token rows are zero
        assert torch.equal(rows.view(torch.int16), _bf16(z[f"rows{step}"]).view(torch.int16))
        full = torch.cat([full, full.new_ones(B, 1)], dim=-1)


def test_tp_sharding_rule_bit_exact_vs_reference_sharder(golden_dir):
    """tests/golden/tp_shard.npz holds what the reference's own sharder (tensor_parallel.py:30-141, lifted and run by
    oracle/gen_golden.py) cuts out of one layer for world sizes 2 and 4.  The oracle's restatement and the product's
    slab packer must cut exactly the same rows / columns; row-parallel biases are the reference's ``bias / world_size``
    (tensor_parallel.py:134 for fc2; LIA applies the same to out_proj, decoder.py:21)."""
    import lia_b200  # noqa: F401
    from lia_b200 import weights
    z = np.load(os.path.join(golden_dir, "tp_shard.npz"))
    h, H, f = int(z["h"]), int(z["H"]), int(z["f"])
    w = {k[2:]: _bf16(z[k]) for k in z.files if k.startswith("w_")}
    for world in (2, 4):
        for rank in range(world):
            tag = f"w{world}r{rank}_"
            ref = {k[len(tag):]: z[k] for k in z.files if k.startswith(tag)}
            assert list(ref["cols"]) == [i * (f // world) for i in range(world + 1)]       # equal 64-blocks per rank
            o = opt_ref.shard_layer(w, H, rank, world)
            p = weights.shard_layer(w, rank, world)
            for k in ("q_w", "q_b", "k_w", "k_b", "v_w", "v_b", "o_w", "fc1_w", "fc1_b", "fc2_w"):
                want = _bf16(ref[k]).view(torch.int16)
                assert torch.equal(o[k].contiguous().view(torch.int16), want), (world, rank, k)
                assert torch.equal(p[k].contiguous().view(torch.int16), want), (world, rank, k)
            # the product pre-divides the row-parallel biases (every rank adds its share before the reduction)
            want_b = _bf16(ref["fc2_b"]).view(torch.int16)
            assert torch.equal(p["fc2_b"].view(torch.int16), want_b), (world, rank)
            assert torch.equal(p["o_b"].view(torch.int16), (w["o_b"].float() / world).to(torch.bfloat16).view(torch.int16))
            # ... and the packed slab holds exactly these shards
            lay = weights.LayerLayout(h, f, world, heads=H)
            v = lay.views(weights.pack_layer(w, lay, rank))
            hq = h // world
            v0 = weights.LayerLayout(h, f, world).views(weights.pack_layer(w, weights.LayerLayout(h, f, world), rank))   # unpadded heads
            assert torch.equal(v0["qkv_w"][hq:2 * hq].view(torch.int16), _bf16(ref["k_w"]).view(torch.int16))
            assert torch.equal(v0["qkv_b"][2 * hq:].view(torch.int16), _bf16(ref["v_b"]).view(torch.int16))
            assert torch.equal(v0["o_w"].view(torch.int16), _bf16(ref["o_w"]).view(torch.int16))
            assert torch.equal(v["fc2_w"].view(torch.int16), _bf16(ref["fc2_w"]).view(torch.int16))
            assert torch.equal(v["fc2_b"].view(torch.int16), want_b)
