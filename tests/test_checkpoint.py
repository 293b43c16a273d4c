"""Weight file formats either side of the path, no GPU needed: the HF checkpoint
forms the reference loads with from_pretrained (run_generation.py:159-167) and its dummy-weight
generator writes (utils/opt-weight-gen.py:66-69, ``safe_serialization=False``), and this build's
native per-layer slab format."""
import json
import os
import struct
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lia_b200  # noqa: E402,F401
from lia_b200 import checkpoint, weights  # noqa: E402
from lia_b200.modeling_opt import OPTConfig  # noqa: E402

BF16 = torch.bfloat16


def tiny_hf_model(dtype=torch.bfloat16):
    from transformers import OPTConfig as HFConfig, OPTForCausalLM as HFOPT
    torch.manual_seed(0)
    c = HFConfig(vocab_size=96, hidden_size=64, num_hidden_layers=3, num_attention_heads=1, ffn_dim=128,
                 max_position_embeddings=32, word_embed_proj_dim=64, do_layer_norm_before=True)
    m = HFOPT(c).eval()
    with torch.no_grad():
        for p in m.parameters():                     # biases / LayerNorm affine away from their 0/1 defaults
            p.add_(torch.randn_like(p) * 0.02)
    return m.to(dtype)


def expected_slabs(m, world):
    sd = {k: v for k, v in m.state_dict().items()}
    cfg = m.config
    lay = weights.LayerLayout(cfg.hidden_size, cfg.ffn_dim, world)
    out = []
    for i in range(cfg.num_hidden_layers):
        w = {k: t.to(BF16) for k, t in weights.layer_from_hf_state_dict(sd, i).items()}
        out.append([weights.pack_layer(w, lay, r) for r in range(world)])
    return lay, out


@pytest.mark.parametrize("form", ["safetensors", "safetensors-sharded", "bin", "bin-sharded"])
def test_hf_checkpoint_forms_give_identical_slabs(tmp_path, form):
    m = tiny_hf_model()
    kw = {"safe_serialization": form.startswith("safetensors")}
    if form.endswith("sharded"):
        kw["max_shard_size"] = "60KB"
    m.save_pretrained(str(tmp_path), **kw)
    names = os.listdir(tmp_path)
    if form.endswith("sharded"):
        assert any(n.endswith(".index.json") for n in names), names
    ck = checkpoint.open_checkpoint(str(tmp_path))
    assert isinstance(ck, checkpoint.HFCheckpoint)
    c = ck.config
    assert (c.hidden_size, c.num_hidden_layers, c.num_attention_heads, c.ffn_dim, c.vocab_size, c.max_position_embeddings) == \
        (64, 3, 1, 128, 96, 32)
    lay, want = expected_slabs(m, 2)
    for i in range(3):
        w = ck.layer(i)
        assert set(w) == set(weights.LAYER_KEYS) and all(t.dtype == BF16 for t in w.values())
        for r in range(2):
            assert torch.equal(weights.pack_layer(w, lay, r), want[i][r])
    e = ck.embeddings()
    sd = m.state_dict()
    assert torch.equal(e["embed_tokens"], sd["model.decoder.embed_tokens.weight"])
    assert torch.equal(e["embed_positions"], sd["model.decoder.embed_positions.weight"]) and e["embed_positions"].shape[0] == 34
    assert torch.equal(e["final_ln_b"], sd["model.decoder.final_layer_norm.bias"])


def test_fp16_checkpoint_and_hub_key_prefix(tmp_path):
    """facebook/opt-* hub files are fp16 and name tensors ``decoder.*`` (no ``model.``)."""
    m = tiny_hf_model(torch.float16)
    sd = {k[len("model."):]: v for k, v in m.state_dict().items() if k.startswith("model.")}
    checkpoint.write_safetensors(str(tmp_path / "model.safetensors"), sd, metadata={"format": "pt"})
    json.dump(m.config.to_dict(), open(tmp_path / "config.json", "w"))
    ck = checkpoint.open_checkpoint(str(tmp_path))
    w = ck.layer(1)
    ref = m.state_dict()["model.decoder.layers.1.fc1.weight"]
    assert w["fc1_w"].dtype == BF16 and torch.equal(w["fc1_w"], ref.to(BF16))


def test_safetensors_reader_against_the_library_and_corruption(tmp_path):
    t = {"a": torch.arange(12, dtype=torch.float32).view(3, 4), "b": torch.randn(5, 7).to(BF16), "s": torch.tensor(3, dtype=torch.int64),
         "e": torch.zeros(0, 4, dtype=torch.float16)}
    p = str(tmp_path / "x.safetensors")
    checkpoint.write_safetensors(p, t, metadata={"k": "v"})
    f = checkpoint.SafetensorsFile(p)
    assert f.metadata == {"k": "v"} and f.data_start % 8 == 0
    for k, v in t.items():
        got = f.get(k)
        assert got.dtype == v.dtype and got.shape == v.shape and torch.equal(got, v)
    try:                                              # the real library reads our writer's output and vice versa
        from safetensors.torch import load_file, save_file
    except ImportError:
        load_file = None
    if load_file is not None:
        back = load_file(p)
        assert all(torch.equal(back[k], v) for k, v in t.items())
        p2 = str(tmp_path / "y.safetensors")
        save_file({k: v for k, v in t.items()}, p2)
        f2 = checkpoint.SafetensorsFile(p2)
        assert all(torch.equal(f2.get(k), v) for k, v in t.items())
    raw = open(p, "rb").read()
    bad = str(tmp_path / "bad.safetensors")
    open(bad, "wb").write(struct.pack("<Q", 1 << 40) + raw[8:])
    with pytest.raises(checkpoint.CheckpointError, match="header length"):
        checkpoint.SafetensorsFile(bad)
    open(bad, "wb").write(raw[:-3])                                   # truncated data section
    with pytest.raises(checkpoint.CheckpointError, match="inconsistent offsets"):
        checkpoint.SafetensorsFile(bad)
    open(bad, "wb").write(raw[:4])
    with pytest.raises(checkpoint.CheckpointError, match="truncated"):
        checkpoint.SafetensorsFile(bad)


@pytest.mark.parametrize("world", [1, 2])
def test_native_slab_round_trip(tmp_path, world):
    m = tiny_hf_model()
    src, dst = tmp_path / "hf", tmp_path / "slabs"
    m.save_pretrained(str(src), safe_serialization=True)
    meta = checkpoint.convert(str(src), str(dst), tp_world=world)
    assert meta["format"] == checkpoint.SLAB_FORMAT and meta["tp_world"] == world
    ck = checkpoint.open_checkpoint(str(dst))
    assert isinstance(ck, checkpoint.SlabCheckpoint) and ck.tp_world == world
    lay, want = expected_slabs(m, world)
    assert os.path.getsize(dst / "rank0.slabs") == lay.nbytes * 3
    out = torch.empty(lay.numel, dtype=BF16)
    for i in (2, 0, 1):                               # random access
        for r in range(world):
            assert torch.equal(ck.read_slab(i, r, out), want[i][r])
    e = ck.embeddings()
    assert torch.equal(e["embed_tokens"], m.state_dict()["model.decoder.embed_tokens.weight"])
    assert torch.equal(e["final_ln_w"], m.state_dict()["model.decoder.final_layer_norm.weight"])
    with pytest.raises(checkpoint.CheckpointError):
        ck.read_slab(0, 0, torch.empty(lay.numel + 8, dtype=BF16))
    with pytest.raises(IndexError):
        ck.read_slab(3, 0, out)
    ck.close()
    with open(dst / "rank0.slabs", "ab") as f:         # a damaged file is refused at open
        f.write(b"\0" * 16)
    with pytest.raises(checkpoint.CheckpointError, match="missing or not"):
        checkpoint.open_checkpoint(str(dst))


def test_unsupported_architectures_are_refused(tmp_path):
    base = dict(model_type="opt", hidden_size=64, num_hidden_layers=1, num_attention_heads=1, ffn_dim=128, vocab_size=96,
                max_position_embeddings=32)
    for extra, exc in (({"_remove_final_layer_norm": True}, NotImplementedError),
                       ({"activation_function": "gelu"}, NotImplementedError), ({"model_type": "llama"}, checkpoint.CheckpointError)):
        json.dump({**base, **extra}, open(tmp_path / "config.json", "w"))
        with pytest.raises(exc):
            checkpoint.config_from_json(str(tmp_path / "config.json"))
    with pytest.raises(checkpoint.CheckpointError, match="not a directory"):
        checkpoint.open_checkpoint(str(tmp_path / "nope"))
    json.dump(base, open(tmp_path / "config.json", "w"))
    with pytest.raises(checkpoint.CheckpointError, match="no model.safetensors"):
        checkpoint.open_checkpoint(str(tmp_path))


def test_opt350m_shaped_checkpoint_round_trip(tmp_path):
    """opt-350m's architecture (LayerNorm after the residual adds, no final LayerNorm, project_in/out around a narrower
    token table; lia/modeling_opt.py:985-1006): HF directory -> config + embeddings -> native slabs -> same tensors."""
    from transformers import OPTConfig as HFConfig, OPTForCausalLM as HFOPT
    torch.manual_seed(1)
    c = HFConfig(vocab_size=96, hidden_size=64, num_hidden_layers=2, num_attention_heads=1, ffn_dim=128,
                 max_position_embeddings=32, word_embed_proj_dim=32, do_layer_norm_before=False)
    m = HFOPT(c).eval().to(BF16)
    src, dst = tmp_path / "hf", tmp_path / "slabs"
    m.save_pretrained(str(src), safe_serialization=True)
    ck = checkpoint.open_checkpoint(str(src))
    assert ck.config.word_embed_proj_dim == 32 and ck.config.embed_dim == 32 and ck.config.do_layer_norm_before is False
    e = ck.embeddings()
    assert set(e) == {"embed_tokens", "embed_positions", "project_in", "project_out"}
    assert e["embed_tokens"].shape == (96, 32) and e["project_in"].shape == (64, 32) and e["project_out"].shape == (32, 64)
    checkpoint.convert(str(src), str(dst))
    nk = checkpoint.open_checkpoint(str(dst))
    assert nk.config.word_embed_proj_dim == 32 and nk.config.do_layer_norm_before is False
    e2 = nk.embeddings()
    assert set(e2) == set(e) and all(torch.equal(e[k], e2[k]) for k in e)
    nk.close()
    # a pre-LN model with equal widths keeps word_embed_proj_dim == 0 ("same as hidden_size")
    tiny_hf_model().save_pretrained(str(tmp_path / "plain"), safe_serialization=True)
    assert checkpoint.open_checkpoint(str(tmp_path / "plain")).config.word_embed_proj_dim == 0


def test_weight_gen_script_writes_dummy_slabs(tmp_path):
    """scripts/opt_weight_gen.py = utils/opt-weight-gen.py without materialising the model: U[0,1) bf16."""
    d = str(tmp_path / "w")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "opt_weight_gen.py"), "--model", "opt-125m", "--save_dir", d,
                        "--tp", "2", "--num-layers", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert "Model saved to" in r.stdout                               # opt-weight-gen.py:67
    ck = checkpoint.open_checkpoint(d)
    assert ck.tp_world == 2 and ck.config.num_hidden_layers == 2 and ck.config.hidden_size == 768
    out = torch.empty(ck.layout.numel, dtype=BF16)
    v = ck.layout.views(ck.read_slab(1, 1, out))
    x = v["fc1_w"].float()
    assert 0.0 <= x.min() and x.max() <= 1.0 and abs(x.mean().item() - 0.5) < 0.01
    full = weights.random_layer(768, 3072, 1000 + 1, "cpu", "dummy")
    assert torch.equal(v["fc1_w"], full["fc1_w"][1536:])
    assert torch.equal(v["o_b"], (full["o_b"].float() / 2).to(BF16))          # row-parallel bias / world
