"""The decode step as ONE persistent kernel (lia_program_*, csrc/decode_program_sm100.cu) against the kernel-per-operation
path it replaces.  Every operation of a program runs the stand-alone kernel's own code in the same order, so the two
paths must agree BIT FOR BIT: tokens, logits, and every layer's K/V cache rows -- for pre-LN and post-LN layers, both head
dims, padded heads, projected embeddings, batch sizes that select each kernel variant (16 / 32 / 64 / 128 rows), and
repeated launches (the dependency counters and stream-K flags carry over from launch to launch).

Reference semantics of the step: lia/modeling_opt.py:1379-1491, decoder.py:172-335, attentions.py:312-557,
models.py:423-431, greedy_search.py:367-395 -- checked against the oracle by the other GPU files through the
kernel-per-operation path; this file pins the program to that path."""
import pytest
import torch

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16


@pytest.fixture(scope="module")
def lia():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import lia_b200
    return lia_b200


def _run(lia, cfg, B, S, new, monkeypatch, program, seed=3, **init):
    from lia_b200 import ops
    monkeypatch.setenv("LIA_DECODE_PROGRAM", "1" if program else "0")
    if not program:
        # the program's attention is the one-CTA-per-(b,h) form (exact rounding points); pin the stand-alone kernel to it
        real = ops.attn_decode
        monkeypatch.setattr(ops, "attn_decode", lambda *a, **k: real(*a, **{**k, "splits": 1}))
    m = lia.OPTForCausalLM(cfg, "cuda").init_weights(seed=seed, bias_std=0.02, ln_std=0.05, **init)
    ids = torch.randint(3, cfg.vocab_size, (B, S), generator=torch.Generator().manual_seed(11))
    outs = [m.generate(ids, max_new_tokens=new, min_new_tokens=new) for _ in range(3)]     # eager / capture / replay, or 3 x program
    st = next(iter(m._states.values()))
    assert bool(st.program) == program, "the program path was not taken" if program else "the program path was taken"
    if program:
        assert st.program.num_ops >= 7 * cfg.num_hidden_layers + 3
    torch.cuda.synchronize()
    res = {"tokens": [o.cpu() for o in outs], "logits": st.logits.clone(), "kc": [k.clone() for k in st.kc], "vc": [v.clone() for v in st.vc],
           "xd": st.xd.clone()}
    for s_ in m._states.values():
        s_.close()
    m._states.clear()
    return res


CASES = [
    # name, cfg kwargs, B, S, new
    ("tiny-d128", dict(hidden_size=256, num_hidden_layers=2, num_attention_heads=2, ffn_dim=1024, vocab_size=512, max_position_embeddings=96), 4, 12, 6),
    ("tiny-d64-b16", dict(hidden_size=256, num_hidden_layers=3, num_attention_heads=4, ffn_dim=512, vocab_size=1000, max_position_embeddings=128), 16, 40, 8),
    ("b32", dict(hidden_size=512, num_hidden_layers=2, num_attention_heads=8, ffn_dim=2048, vocab_size=2048, max_position_embeddings=300), 32, 200, 5),
    ("b64-long", dict(hidden_size=1024, num_hidden_layers=2, num_attention_heads=8, ffn_dim=4096, vocab_size=4096, max_position_embeddings=400), 64, 300, 6),
    ("b128", dict(hidden_size=512, num_hidden_layers=2, num_attention_heads=4, ffn_dim=1024, vocab_size=1024, max_position_embeddings=64), 128, 20, 5),
    ("b1", dict(hidden_size=256, num_hidden_layers=2, num_attention_heads=2, ffn_dim=512, vocab_size=512, max_position_embeddings=64), 1, 9, 5),
    ("postln-projected", dict(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, ffn_dim=1024, vocab_size=512, max_position_embeddings=64,
                              do_layer_norm_before=False, word_embed_proj_dim=128), 4, 10, 6),
    ("headdim80", dict(hidden_size=320, num_hidden_layers=2, num_attention_heads=4, ffn_dim=640, vocab_size=512, max_position_embeddings=64), 8, 12, 5),
    ("h7168-row", dict(hidden_size=7168, num_hidden_layers=1, num_attention_heads=56, ffn_dim=28672, vocab_size=50272, max_position_embeddings=64), 64, 16, 4),
]


@pytest.mark.parametrize("name,kw,B,S,new", CASES, ids=[c[0] for c in CASES])
def test_program_is_bit_identical_to_kernel_per_operation(lia, monkeypatch, name, kw, B, S, new):
    cfg = lia.OPTConfig(**kw)
    ref = _run(lia, cfg, B, S, new, monkeypatch, program=False)
    got = _run(lia, cfg, B, S, new, monkeypatch, program=True)
    assert torch.equal(ref["tokens"][0], ref["tokens"][2])
    for i in range(3):
        assert torch.equal(got["tokens"][i], ref["tokens"][0]), f"{name}: tokens of program run {i} differ"
    assert torch.equal(got["logits"], ref["logits"]), f"{name}: last-step logits differ"
    assert torch.equal(got["xd"], ref["xd"]), f"{name}: last hidden state differs"
    for li, (a, b) in enumerate(zip(got["kc"], ref["kc"])):
        assert torch.equal(a, b), f"{name}: K cache of layer {li} differs"
    for li, (a, b) in enumerate(zip(got["vc"], ref["vc"])):
        assert torch.equal(a, b), f"{name}: V cache of layer {li} differs"


def test_program_with_padded_prompts_and_eos_suppression(lia, monkeypatch):
    """The mask-driven learned positions (lia/modeling_opt.py:368-378) and the min_new_tokens eos suppression
    (generation_utils.py:872-880) are launch arguments / operations of the program too."""
    cfg = lia.OPTConfig(hidden_size=256, num_hidden_layers=2, num_attention_heads=2, ffn_dim=512, vocab_size=64, max_position_embeddings=64)
    ids = torch.randint(3, cfg.vocab_size, (6, 14), generator=torch.Generator().manual_seed(5))
    ids[0, :4] = cfg.pad_token_id
    ids[3, :9] = cfg.pad_token_id
    outs = {}
    for program in (False, True):
        monkeypatch.setenv("LIA_DECODE_PROGRAM", "1" if program else "0")
        m = lia.OPTForCausalLM(cfg, "cuda").init_weights(seed=9, bias_std=0.05, ln_std=0.1)
        outs[program] = [m.generate(ids, max_new_tokens=10, min_new_tokens=mn).cpu() for mn in (10, 4, 0)]
        st = next(iter(m._states.values()))
        assert bool(st.program) == program
        for s_ in m._states.values():
            s_.close()
    for a, b in zip(outs[True], outs[False]):
        assert torch.equal(a, b)
