"""Host-side logic that needs no GPU: slab layout, TP sharding, the stream-K partition the decode
GEMM uses, config table, CLI flags, policy mapping, the CPU-baseline arm of bench.py."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lia_b200  # noqa: E402
from lia_b200 import weights  # noqa: E402
from lia_b200.modeling_opt import OPT_CONFIGS, _check_policy, get_config  # noqa: E402
from oracle import opt_ref  # noqa: E402


def test_config_table_matches_survey():
    dims = {"opt-1.3b": (24, 2048, 32, 8192), "opt-30b": (48, 7168, 56, 28672), "opt-66b": (64, 9216, 72, 36864),
            "opt-175b": (96, 12288, 96, 49152)}
    for k, (L, h, H, f) in dims.items():
        c = get_config("facebook/" + k)
        assert (c.num_hidden_layers, c.hidden_size, c.num_attention_heads, c.ffn_dim) == (L, h, H, f)
        assert c.vocab_size == 50272 and c.max_position_embeddings == 2048 and c.head_dim in (64, 128)
        lay = weights.LayerLayout(h, f)
        assert lay.nbytes == 2 * (12 * h * h + 13 * h)           # W_l of SURVEY.md 8d
    c = get_config("opt-30b")
    c.num_hidden_layers = 2
    assert OPT_CONFIGS["opt-30b"].num_hidden_layers == 48        # get_config returns a copy


@pytest.mark.parametrize("world", [1, 2, 4])
def test_slab_pack_and_tp_sharding(world):
    h, f = 256, 1024
    w = weights.random_layer(h, f, seed=3, bias_std=0.05, ln_std=0.1)
    assert torch.equal(w["q_w"], weights.random_layer(h, f, seed=3, bias_std=0.05, ln_std=0.1)["q_w"])
    lay = weights.LayerLayout(h, f, world)
    for off, n in lay.offsets.values():
        assert off % 8 == 0 and n % 8 == 0                        # 16-byte aligned tensors
    x = torch.randn(5, h).to(torch.bfloat16)
    full = torch.relu(x.float() @ w["fc1_w"].float().t() + w["fc1_b"].float()) @ w["fc2_w"].float().t() + w["fc2_b"].float()
    total = 0
    for r in range(world):
        v = lay.views(weights.pack_layer(w, lay, r))
        ref = opt_ref.shard_layer(w, 4, r, world)                  # oracle's restatement of tensor_parallel.py
        hq = h // world
        assert torch.equal(v["qkv_w"][:hq], ref["q_w"]) and torch.equal(v["qkv_w"][2 * hq:], ref["v_w"])
        assert torch.equal(v["o_w"], ref["o_w"]) and torch.equal(v["fc1_w"], ref["fc1_w"]) and torch.equal(v["fc2_w"], ref["fc2_w"])
        assert torch.equal(v["qkv_b"][hq:2 * hq], ref["k_b"])
        # row-parallel bias is divided by the world size (tensor_parallel.py:134)
        assert torch.allclose(v["fc2_b"].float() * world, w["fc2_b"].float(), atol=1e-2)
        total = total + torch.relu(x.float() @ v["fc1_w"].float().t() + v["fc1_b"].float()) @ v["fc2_w"].float().t() + v["fc2_b"].float()
    assert torch.allclose(total, full, atol=2e-2, rtol=2e-2)


@pytest.mark.parametrize("H,d,world", [(4, 80, 1), (4, 80, 2), (8, 32, 4), (2, 96, 1), (4, 64, 2), (2, 128, 2), (6, 40, 3)])
def test_head_padding_is_exact_and_invertible(H, d, world):
    """Zero-padded heads (weights.padded_head_dim): per rank, the padded slab holds exactly the shard's values in the
    first d lanes of every head and zeros elsewhere, and the padded projections compute the unpadded ones:
    (x Wq^T + bq) in lanes [:d], exact zeros in the rest; ctx_pad Wo_pad^T == ctx Wo^T."""
    h, f = H * d, 2 * H * d
    dp = 64 if d <= 64 else 128
    w = weights.random_layer(h, f, seed=7, bias_std=0.05, ln_std=0.1)
    for r in range(world):
        lay = weights.LayerLayout(h, f, world, heads=H)
        Hl = H // world
        assert (lay.d, lay.dp, lay.heads_local, lay.hq) == (d, dp, Hl, Hl * dp)
        v = lay.views(weights.pack_layer(w, lay, r))
        ref = opt_ref.shard_layer(w, H, r, world)
        for i, n in enumerate(("q", "k", "v")):
            blk = v["qkv_w"][i * lay.hq:(i + 1) * lay.hq].view(Hl, dp, h)
            assert torch.equal(blk[:, :d].reshape(Hl * d, h), ref[n + "_w"]) and not blk[:, d:].any()
            bb = v["qkv_b"][i * lay.hq:(i + 1) * lay.hq].view(Hl, dp)
            assert torch.equal(bb[:, :d].reshape(-1), ref[n + "_b"]) and not bb[:, d:].any()
        ow = v["o_w"].view(h, Hl, dp)
        assert torch.equal(ow[:, :, :d].reshape(h, Hl * d), ref["o_w"]) and not ow[:, :, d:].any()
        x = torch.randn(5, h, dtype=torch.float64)
        qp = (x @ v["qkv_w"][:lay.hq].double().t() + v["qkv_b"][:lay.hq].double()).view(5, Hl, dp)
        qr = (x @ ref["q_w"].double().t() + ref["q_b"].double()).view(5, Hl, d)
        assert torch.equal(qp[..., :d], qr) and not qp[..., d:].any()
        ctx = torch.randn(5, Hl, d, dtype=torch.float64)
        ctx_pad = torch.zeros(5, Hl, dp, dtype=torch.float64)
        ctx_pad[..., :d] = ctx
        assert torch.allclose(ctx_pad.view(5, -1) @ v["o_w"].double().t(), ctx.view(5, -1) @ ref["o_w"].double().t(), atol=1e-12)
    if dp == d:                       # no padding: the layout is byte-for-byte the unpadded one
        assert weights.LayerLayout(h, f, world, heads=H).offsets == weights.LayerLayout(h, f, world).offsets
    with pytest.raises(NotImplementedError):
        weights.LayerLayout(4 * 136, 64, 1, heads=4)
    with pytest.raises(NotImplementedError):
        weights.LayerLayout(4 * 20, 64, 1, heads=4)


def streamk_spans(tiles, k_blocks, grid):
    """Python restatement of Sched<SWAP> in csrc/gemm_sm100.cu."""
    total = tiles * k_blocks
    out = []
    for c in range(grid):
        pos, end = total * c // grid, total * (c + 1) // grid
        segs = []
        while pos < end:
            t = pos // k_blocks
            kb0 = pos - t * k_blocks
            kb1 = min(k_blocks, kb0 + (end - pos))
            segs.append((t, kb0, kb1))
            pos += kb1 - kb0
        out.append(segs)
    return out


@pytest.mark.parametrize("tiles,k_blocks,grid", [(56, 112, 148), (168, 112, 148), (224, 112, 148), (56, 448, 148),
                                                 (393, 112, 148), (1, 2, 1), (3, 5, 3), (2, 1, 1), (7, 33, 148 // 4)])
def test_streamk_partition_properties(tiles, k_blocks, grid):
    spans = streamk_spans(tiles, k_blocks, grid)
    cover = {}
    for c, segs in enumerate(spans):
        non_owner = [s for s in segs if s[1] > 0]
        partial_owner = [s for s in segs if s[1] == 0 and s[2] < k_blocks]
        assert len(non_owner) <= 1 and len(partial_owner) <= 1           # one workspace slot per CTA suffices
        if non_owner:
            assert segs[0] == non_owner[0]                               # published first: the owner never deadlocks
        if partial_owner:
            assert segs[-1] == partial_owner[0]
        for (t, a, b) in segs:
            for kb in range(a, b):
                assert (t, kb) not in cover
                cover[(t, kb)] = c
    assert len(cover) == tiles * k_blocks                                # every k-block exactly once
    sizes = [sum(b - a for _, a, b in segs) for segs in spans]
    assert max(sizes) - min(sizes) <= 1                                  # every SM streams the same bytes
    # the owner of a split tile finds its contributors as the consecutive CTAs c+1.. (kernel's loop)
    for c, segs in enumerate(spans):
        for (t, a, b) in segs:
            if a == 0 and b < k_blocks:
                last = c
                while last + 1 < grid and (tiles * k_blocks) * (last + 1) // grid < (t + 1) * k_blocks:
                    last += 1
                contributors = sorted({cover[(t, kb)] for kb in range(b, k_blocks)})
                assert contributors == list(range(c + 1, last + 1))


def test_policy_mapping():
    for p in (0, 2, 3, 4):
        assert _check_policy(p, "prefill_policy") == p
    assert _check_policy(None, "x") == 3
    with pytest.raises(NotImplementedError):
        _check_policy(1, "decoding_policy")
    with pytest.raises(ValueError):
        _check_policy(7, "decoding_policy")


def test_cli_has_every_reference_flag():
    from lia_b200.run import build_parser, main
    flags = {a for act in build_parser()._actions for a in act.option_strings}
    # run.py:169-215 / run_generation.py:59-118
    for f in ["-m", "--dtype", "--ipex", "--benchmark", "--input-tokens", "--max-new-tokens", "--batch-size", "--num-iter",
              "--num-warmup", "--greedy", "--token-latency", "--profile", "--prefill-policy", "--decoding-policy",
              "--no-overlap", "--pin-weight", "--gpu-percentage", "--num-minibatch", "--enable-cxl"]:
        assert f in flags, f
    a = build_parser().parse_args([])
    assert (a.prefill_policy, a.decoding_policy, a.gpu_percentage, a.num_minibatch, a.max_new_tokens, a.num_iter,
            a.num_warmup) == (1, 1, 0, 1, 32, 100, 10)                    # reference defaults
    assert main(["-m", "facebook/opt-1.3b"]) == 2                        # default policy 1/1 = CPU path: refused


def test_cli_prompt_and_tokenizer_inputs(tmp_path, capsys):
    """Caller side of the path (run_generation.py:169-171, 260-285, 319): a checkpoint directory that carries a tokenizer
    turns --prompt / prompt.json text into ids (every batch row the same text) and ids back into text; without one the
    ids are synthetic.  No GPU involved: only the input builder is exercised."""
    from tokenizers import Tokenizer, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast
    from lia_b200.run import build_inputs, build_parser, load_tokenizer, pick_prompt
    words = ["<pad>", "</s>", "<unk>"] + "the quick brown fox jumps over a lazy dog and keeps running".split()
    tk = Tokenizer(models.WordLevel({w: i for i, w in enumerate(words)}, unk_token="<unk>"))
    tk.pre_tokenizer = pre_tokenizers.Whitespace()
    d = tmp_path / "ckpt"
    d.mkdir()
    PreTrainedTokenizerFast(tokenizer_object=tk, pad_token="<pad>", eos_token="</s>", unk_token="<unk>").save_pretrained(str(d))
    tok = load_tokenizer(str(d))
    assert tok is not None and load_tokenizer("facebook/opt-30b") is None and load_tokenizer(str(tmp_path)) is None
    args = build_parser().parse_args(["--batch-size", "3", "--input-tokens", "6", "--prompt", "the quick brown fox"])
    ids, text = build_inputs(args, 50272, tok)
    assert text == "the quick brown fox" and ids.shape == (3, 4) and ids.dtype == torch.int64
    assert torch.equal(ids[0], ids[2]) and ids[0].tolist() == [3, 4, 5, 6]
    assert "---- Prompt size: 4" in capsys.readouterr().out                       # run_generation.py:279
    assert tok.batch_decode(ids, skip_special_tokens=True)[1] == "the quick brown fox"
    # prompt.json pool keyed by model type and --input-tokens (run_generation.py:262-276)
    json.dump({"opt": {"6": "a lazy dog keeps running and"}}, open(tmp_path / "prompt.json", "w"))
    args = build_parser().parse_args(["--batch-size", "2", "--input-tokens", "6"])
    assert pick_prompt(args, [str(tmp_path)]) == "a lazy dog keeps running and"
    ids, text = build_inputs(args, 50272, tok, [str(tmp_path)])
    assert ids.shape == (2, 6) and text.startswith("a lazy dog")
    # no tokenizer (or no prompt for that length): synthetic ids of exactly --input-tokens, seeded, never pad/eos
    ids, text = build_inputs(args, 50272, None)
    assert text is None and ids.shape == (2, 6) and int(ids.min()) >= 3 and torch.equal(ids, build_inputs(args, 50272, None)[0])
    args = build_parser().parse_args(["--input-tokens", "7"])
    assert build_inputs(args, 50272, tok, [str(tmp_path)])[1] is None


def test_bench_reference_arm_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--model", "opt-125m",
                          "--batch-size", "2", "--input-tokens", "16", "--max-new-tokens", "4", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "tokens/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["higher_is_better"] is True


def test_layer_streamer_call_sequences(monkeypatch):
    """LayerStreamer's calls into the C ABI (lia_streamer_prefetch / wait / release), recorded against a fake library:
    the double-buffered schedule (layer j in slot j % 2, layer j+2 prefetched when j is released, wrapping into the next
    forward) and the --no-overlap schedule (everything through slot 0, the copy ordered after all earlier compute by a
    release recorded at acquire time, nothing fetched ahead)."""
    from lia_b200 import _lib, streamer

    class Fake:
        def __init__(self):
            self.log = []

        def lia_streamer_create(self, arr, n, nbytes):
            return 1

        def lia_streamer_prefetch(self, h, slot, ptr, nbytes):
            self.log.append(("prefetch", slot, ptr))
            return 0

        def lia_streamer_wait(self, h, slot, stream):
            self.log.append(("wait", slot))
            return 0

        def lia_streamer_release(self, h, slot, stream):
            self.log.append(("release", slot))
            return 0

        def lia_streamer_destroy(self, h):
            return 0
    fake = Fake()
    monkeypatch.setattr(_lib, "load", lambda: fake)

    class S:
        cuda_stream = 0
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a: S())
    lay = weights.LayerLayout(64, 128)
    host = [torch.zeros(lay.numel, dtype=torch.bfloat16) for _ in range(4)]
    ptr = [t.data_ptr() for t in host]
    st = streamer.LayerStreamer(lay, host, "cpu")

    def forward():
        fake.log.clear()
        st.begin()
        for j in range(4):
            st.acquire(j)
            st.release(j)
        return list(fake.log)

    first = forward()
    assert first == [("prefetch", 0, ptr[0]), ("prefetch", 1, ptr[1]),
                     ("wait", 0), ("release", 0), ("prefetch", 0, ptr[2]),
                     ("wait", 1), ("release", 1), ("prefetch", 1, ptr[3]),
                     ("wait", 0), ("release", 0), ("prefetch", 0, ptr[0]),       # wraps into the next forward
                     ("wait", 1), ("release", 1), ("prefetch", 1, ptr[1])]
    second = forward()                       # the first two layers are already in flight: no prefetch in begin()
    assert second == first[2:]
    st.overlap = False
    serial = forward()
    # slot 0 holds layer 0 already (wrap-around prefetch of the previous forward): no copy; every later layer is copied
    # on demand into slot 0, after a release that orders the copy behind all compute enqueued so far
    assert serial == [("wait", 0), ("release", 0),
                      ("release", 0), ("prefetch", 0, ptr[1]), ("wait", 0), ("release", 0),
                      ("release", 0), ("prefetch", 0, ptr[2]), ("wait", 0), ("release", 0),
                      ("release", 0), ("prefetch", 0, ptr[3]), ("wait", 0), ("release", 0)]
    serial2 = forward()                      # slot 0 now holds layer 3: layer 0 is copied again
    assert serial2[:4] == [("release", 0), ("prefetch", 0, ptr[0]), ("wait", 0), ("release", 0)]
    assert all(c[1] == 0 for c in serial + serial2)
    st.overlap = True                        # back to double buffering: both slots are (re)filled as needed
    again = forward()                        # slot 1 still holds layer 1 from the last double-buffered forward
    assert again == [("prefetch", 0, ptr[0])] + first[2:]
    st.close()
