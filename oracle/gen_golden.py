#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from the REFERENCE ITSELF.

TEST INFRASTRUCTURE ONLY.  Runs in the build container (needs /root/reference);
its outputs (small .npz files) are committed so that nothing on the GPU box ever
reads /root/reference.

Two kinds of fixtures are produced:

1. ``layer_*.npz`` -- outputs of the reference's OWN decoder-layer code.  The
   reference package cannot be imported here (IPEX C++
   extension unbuildable, ``deepspeed`` missing), so the two functions that hold
   all the math of the hot path are lifted *at run time* out of the reference
   source files with ``ast`` and executed unmodified except for one textual
   substitution, ``'cuda'`` -> ``'cpu'`` (this container has no GPU):

     * ``OPTDecoderLayer_forward`` and its ``gpu_*`` helpers
       (intel_extension_for_pytorch/transformers/models/reference/modules/decoder.py:18-119,172-335)
     * ``_OPTAttention_forward``
       (intel_extension_for_pytorch/transformers/models/reference/modules/attentions.py:312-557)

   They are driven through the ``policy == 3`` ("everything on GPU, KV on GPU")
   branch exactly as lia/modeling_opt.py:1246-1260 drives a resident layer: one
   prefill call and then decode calls that re-use the returned 4-tuple cache.
   No source text of the reference is copied into this repository.

2. ``model_hf_*.npz`` -- logits and greedy tokens from the stock
   ``transformers.OPTForCausalLM`` (an independent implementation of the same
   model) in fp32, for a whole-model cross-check at the reference's own nightly
   tolerance (tests/cpu/test_ipex_optimize_transformers_nightly.py:237, prec 0.1).

Usage:  python oracle/gen_golden.py            (writes tests/golden/*.npz)
"""
import ast
import os
import sys
import types

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F
from typing import List, Optional, Tuple, Union

REF = "/root/reference"
D_PATH = os.path.join(REF, "intel_extension_for_pytorch/transformers/models/reference/modules/decoder.py")
A_PATH = os.path.join(REF, "intel_extension_for_pytorch/transformers/models/reference/modules/attentions.py")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _lift(path, names):
    """Return {name: function} for top-level functions ``names`` of ``path``."""
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "nn": nn, "F": F, "Optional": Optional, "Tuple": Tuple,
          "Union": Union, "List": List}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            text = ast.get_source_segment(src, node).replace("'cuda'", "'cpu'")
            exec(compile(text, path + ":" + node.name, "exec"), ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing
    return ns


D_FUNCS = ["gpu_mha_linear_load", "gpu_fc1_linear_load", "gpu_fc2_linear_load",
           "gpu_linear_compute", "gpu_linear_compute_no_delete",
           "gpu_linear_relu_compute", "gpu_linear_relu_compute_no_delete",
           "gpu_ln_compute_self_attn", "gpu_ln_compute_final",
           "gpu_mha_linear_load_ds", "gpu_fc1_linear_load_ds", "gpu_fc2_linear_load_ds",
           "gpu_linear_allreduce_compute", "gpu_linear_allreduce_compute_no_delete",
           "OPTDecoderLayer_forward"]


def bf16_bits(t):
    return t.detach().contiguous().view(torch.int16).numpy().view(np.uint16).copy()


def make_layer_weights(h, f, seed, kind="normal"):
    """normal(0, 0.02) weights; small non-zero biases and non-trivial LN affine so
    that every term of the path is exercised (lia/modeling_opt.py:895-904 uses
    zero biases; zeros would hide bias/LN-affine bugs)."""
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std=0.02, mean=0.0):
        return (torch.randn(*shape, generator=g) * std + mean).to(torch.bfloat16)

    w = {}
    w["ln1_w"] = rn(h, std=0.1, mean=1.0); w["ln1_b"] = rn(h, std=0.1)
    for n in ("q", "k", "v", "o"):
        w[n + "_w"] = rn(h, h); w[n + "_b"] = rn(h, std=0.05)
    w["ln2_w"] = rn(h, std=0.1, mean=1.0); w["ln2_b"] = rn(h, std=0.1)
    w["fc1_w"] = rn(f, h); w["fc1_b"] = rn(f, std=0.05)
    w["fc2_w"] = rn(h, f); w["fc2_b"] = rn(h, std=0.05)
    return w


class _Holder:
    pass


def build_reference_layer(w, h, H, pre_ln=True):
    """Fake ``self`` objects carrying exactly the attributes the two lifted
    functions read on the policy-3 path (names as set up by
    _IPEXDecoderLayerRef.__init__, decoder.py:1385-1392, and by
    move_gpu_layer, lia/modeling_opt.py:254-267)."""
    d = h // H
    attn_ns = _lift(A_PATH, ["_OPTAttention_forward"])
    dec_ns = _lift(D_PATH, D_FUNCS)

    def lin(wt, b):
        m = nn.Linear(wt.shape[1], wt.shape[0], bias=True, dtype=torch.bfloat16)
        m.weight = nn.Parameter(wt.clone(), requires_grad=False)
        m.bias = nn.Parameter(b.clone(), requires_grad=False)
        return m

    def ln(wt, b):
        m = nn.LayerNorm(h, dtype=torch.bfloat16)          # eps 1e-5 (module default)
        m.weight = nn.Parameter(wt.clone(), requires_grad=False)
        m.bias = nn.Parameter(b.clone(), requires_grad=False)
        return m

    attn = _Holder()
    attn.num_heads, attn.head_dim, attn.embed_dim = H, d, h
    attn.scaling = d ** -0.5                                 # lia/modeling_opt.py:413
    attn.is_decoder = True
    attn.q_proj, attn.k_proj, attn.v_proj = lin(w["q_w"], w["q_b"]), lin(w["k_w"], w["k_b"]), lin(w["v_w"], w["v_b"])

    layer = _Holder()
    layer.distributed = False
    layer.do_layer_norm_before = pre_ln                     # False: opt-350m (decoder.py:250-259, 320-321)
    layer.self_attn_layer_norm = ln(w["ln1_w"], w["ln1_b"])
    layer.final_layer_norm = ln(w["ln2_w"], w["ln2_b"])
    layer.mha_linear_add = _Holder(); layer.mha_linear_add.weight = w["o_w"]; layer.mha_linear_add.bias = w["o_b"]
    layer.mlp_linear_add = _Holder(); layer.mlp_linear_add.weight = w["fc2_w"]; layer.mlp_linear_add.bias = w["fc2_b"]
    layer.linear_relu = _Holder(); layer.linear_relu.linear = lin(w["fc1_w"], w["fc1_b"])
    layer.self_attn = lambda **kw: attn_ns["_OPTAttention_forward"](attn, **kw)

    def call(x, past, max_new_tokens):
        with torch.no_grad():
            return dec_ns["OPTDecoderLayer_forward"](
                layer, x, attention_mask=torch.ones(1), layer_head_mask=None,
                output_attentions=False, use_cache=True, past_key_value=past,
                gpu_layer=None, policy=3, max_new_tokens=max_new_tokens)
    return call


def gen_layer_case(name, B, S, h, H, new, seed, pre_ln=True):
    f = 4 * h
    w = make_layer_weights(h, f, seed)
    call = build_reference_layer(w, h, H, pre_ln)
    g = torch.Generator().manual_seed(seed + 1000)
    xs = [torch.randn(B, S, h, generator=g).to(torch.bfloat16)]
    xs += [torch.randn(B, 1, h, generator=g).to(torch.bfloat16) for _ in range(new)]
    # initial fake past: intel_extension_for_pytorch/transformers/generation/greedy_search.py:272-282
    past = (torch.zeros(1, 0, 0, 1, dtype=torch.long), torch.zeros(1, 1, 1, 1), torch.zeros(1, 1, 1, 1),
            torch.zeros(2048, B, dtype=torch.long))
    out = {"B": B, "S": S, "h": h, "H": H, "new": new, "seed": seed, "pre_ln": int(pre_ln)}
    for k, v in w.items():
        out["w_" + k] = bf16_bits(v)
    for step, x in enumerate(xs):
        y, past = call(x, past, new)
        out[f"x{step}"] = bf16_bits(x)
        out[f"y{step}"] = bf16_bits(y)
        assert past[0].shape[2] == (S + step), past[0].shape
    T = S + new
    out["kcache"] = bf16_bits(past[1][:T])                  # [T, B, H, d] time-major
    out["vcache"] = bf16_bits(past[2][:T])
    np.savez_compressed(os.path.join(OUT, name), **out)
    print("wrote", name, "bytes", os.path.getsize(os.path.join(OUT, name + ".npz")))


def gen_hf_model_case(name, V, h, H, L, P, B, S, new, seed, word_dim=None, pre_ln=True):
    from transformers import OPTConfig, OPTForCausalLM
    torch.manual_seed(seed)
    cfg = OPTConfig(vocab_size=V, hidden_size=h, num_attention_heads=H, num_hidden_layers=L,
                    ffn_dim=4 * h, max_position_embeddings=P, word_embed_proj_dim=word_dim or h,
                    do_layer_norm_before=pre_ln, activation_function="relu", dropout=0.0,
                    pad_token_id=1, bos_token_id=2, eos_token_id=2, init_std=0.02)
    model = OPTForCausalLM(cfg).eval().float()
    # make every parameter exactly bf16-representable so that the fixture can be stored as bf16 bits
    with torch.no_grad():
        for p_ in model.parameters():
            p_.copy_(p_.to(torch.bfloat16).float())
        # non-zero biases / LN affine (seeded) so these terms are exercised
        g = torch.Generator().manual_seed(seed + 7)
        for n_, p_ in model.named_parameters():
            if n_.endswith("bias"):
                p_.copy_((torch.randn(p_.shape, generator=g) * 0.05).to(torch.bfloat16).float())
            elif "layer_norm.weight" in n_:
                p_.copy_((1 + torch.randn(p_.shape, generator=g) * 0.1).to(torch.bfloat16).float())
    g = torch.Generator().manual_seed(seed + 1)
    ids = torch.randint(3, V, (B, S), generator=g)
    with torch.no_grad():
        logits = model(ids).logits[:, -1, :]
        toks = model.generate(ids, do_sample=False, num_beams=1, max_new_tokens=new, min_new_tokens=new)
    out = {"V": V, "h": h, "H": H, "L": L, "P": P, "B": B, "S": S, "new": new, "word_dim": word_dim or h,
           "pre_ln": int(pre_ln), "input_ids": ids.numpy(), "prefill_last_logits": logits.numpy().astype(np.float32),
           "tokens": toks.numpy()}
    sd = model.state_dict()
    for k, v in sd.items():
        if k == "lm_head.weight":
            continue                                           # tied (lia/modeling_opt.py:1660)
        out["sd:" + k] = bf16_bits(v.to(torch.bfloat16))
    np.savez_compressed(os.path.join(OUT, name), **out)
    print("wrote", name, "bytes", os.path.getsize(os.path.join(OUT, name + ".npz")))


def _lift_class_method(path, cls, method, ns):
    """Compile one method of a top-level class of ``path`` as a plain function in ``ns``."""
    src = open(path).read()
    for node in ast.parse(src).body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name == method:
                    text = ast.get_source_segment(src, sub)
                    import textwrap
                    exec(compile(textwrap.dedent(text), f"{path}:{cls}.{method}", "exec"), ns)
                    return ns[method]
    raise KeyError((cls, method))


def gen_positions_case(name, seed):
    """The reference's own OPTLearnedPositionalEmbedding.forward (lia/modeling_opt.py:368-378) and
    _prepare_attention_mask_for_generation (lia/generation_utils.py:469-485), lifted at run time, on left-padded,
    right-padded and hole-y masks.  forward() ends in ``super().forward(positions + self.offset)``: it is executed
    with ``super`` bound to a shim whose forward is nn.Embedding's lookup, so the stored rows are the reference's."""
    M_PATH = os.path.join(REF, "lia/modeling_opt.py")
    GU_PATH = os.path.join(REF, "lia/generation_utils.py")
    g = torch.Generator().manual_seed(seed)
    P, h, B, S, new = 40, 16, 6, 9, 4
    table = torch.randn(P + 2, h, generator=g).to(torch.bfloat16)

    class _Super:
        def forward(self_inner, idx):
            _Super.last = idx.clone()
            return F.embedding(idx, table)

    class _Self:
        offset = 2                                                           # lia/modeling_opt.py:365
    ns = {"torch": torch, "nn": nn, "F": F, "Optional": Optional, "Tuple": Tuple, "Union": Union, "List": List,
          "super": lambda *a: _Super()}
    fwd = _lift_class_method(M_PATH, "OPTLearnedPositionalEmbedding", "forward", ns)
    prep = _lift_class_method(GU_PATH, "GenerationMixin", "_prepare_attention_mask_for_generation", dict(ns))
    ids = torch.randint(3, 300, (B, S), generator=g)
    ids[0, :3] = 1                      # left padding
    ids[1, :1] = 1
    ids[2, -2:] = 1                     # right padding
    ids[3, 4] = 1                       # a hole
    ids[4, :] = 1                       # fully padded row
    mask = prep(None, ids, 1, 2)
    assert mask.sum() < mask.numel()
    mask_nopad = prep(None, ids.clamp(min=3), 1, 2)
    mask_pad_is_eos = prep(None, ids, 1, 1)
    out = {"P": P, "h": h, "B": B, "S": S, "new": new, "table": bf16_bits(table), "ids": ids.numpy(),
           "mask": mask.numpy(), "mask_nopad": mask_nopad.numpy(), "mask_pad_is_eos": mask_pad_is_eos.numpy()}
    full = mask
    for step in range(new + 1):
        past = 0 if step == 0 else S + step - 1
        rows = fwd(_Self(), full, past)
        out[f"pos{step}"] = _Super.last.numpy()
        out[f"rows{step}"] = bf16_bits(rows)
        full = torch.cat([full, full.new_ones(B, 1)], dim=-1)                 # greedy_search.py:411
    np.savez_compressed(os.path.join(OUT, name), **out)
    print("wrote", name, "bytes", os.path.getsize(os.path.join(OUT, name + ".npz")))


def gen_tp_shard_case(name, seed):
    """The reference's OWN tensor-parallel sharder (intel_extension_for_pytorch/transformers/tensor_parallel.py:30-141:
    TensorParallellLinear.shard_weights_by_head / shard_weights_by_block, lifted at run time) applied to one layer's
    q/k/v (column split by heads), out_proj (row split by heads), fc1 (column split by 64-blocks) and fc2 (row split by
    64-blocks, bias / world_size :134) for world sizes 2 and 4; stored per rank."""
    TP_PATH = os.path.join(REF, "intel_extension_for_pytorch/transformers/tensor_parallel.py")
    ns = {"torch": torch, "nn": nn}
    by_head = _lift_class_method(TP_PATH, "TensorParallellLinear", "shard_weights_by_head", dict(ns))
    by_block = _lift_class_method(TP_PATH, "TensorParallellLinear", "shard_weights_by_block", dict(ns))
    h, H, f = 64, 4, 256
    d = h // H
    w = make_layer_weights(h, f, seed)

    def lin(wt, b):
        m = nn.Linear(wt.shape[1], wt.shape[0], bias=True, dtype=torch.bfloat16)
        m.weight = nn.Parameter(wt.clone(), requires_grad=False)
        m.bias = nn.Parameter(b.clone(), requires_grad=False)
        return m
    out = {"h": h, "H": H, "f": f, "seed": seed}
    for k, v in w.items():
        out["w_" + k] = bf16_bits(v)
    for world in (2, 4):
        for rank in range(world):
            tag = f"w{world}r{rank}_"
            for n in ("q", "k", "v"):                                   # shard_mha_weights: column split by heads
                wt, b = by_head(None, lin(w[n + "_w"], w[n + "_b"]), H, H, d, rank, world, True)
                out[tag + n + "_w"], out[tag + n + "_b"] = bf16_bits(wt.data), bf16_bits(b.data)
            wt, _ = by_head(None, lin(w["o_w"], w["o_b"]), H, H, d, rank, world, False)       # row split by heads
            out[tag + "o_w"] = bf16_bits(wt.data)
            wt, b, cols = by_block(None, lin(w["fc1_w"], w["fc1_b"]), rank, world, True)      # shard_mlp_weights
            out[tag + "fc1_w"], out[tag + "fc1_b"] = bf16_bits(wt.data), bf16_bits(b.data)
            wt, b, cols = by_block(None, lin(w["fc2_w"], w["fc2_b"]), rank, world, False)
            out[tag + "fc2_w"], out[tag + "fc2_b"] = bf16_bits(wt.data), bf16_bits(b.data)   # bias / world_size
            out[tag + "cols"] = np.array(cols)
    np.savez_compressed(os.path.join(OUT, name), **out)
    print("wrote", name, "bytes", os.path.getsize(os.path.join(OUT, name + ".npz")))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)                                   # bit-stable CPU reductions
    gen_layer_case("layer_d64", B=3, S=8, h=128, H=2, new=3, seed=11)
    gen_layer_case("layer_d128", B=2, S=5, h=128, H=1, new=2, seed=12)
    gen_layer_case("layer_ragged", B=1, S=1 + 16, h=192, H=3, new=1, seed=13)   # h not a power of two
    gen_positions_case("positions_padded", seed=31)
    gen_tp_shard_case("tp_shard", seed=41)
    gen_hf_model_case("model_hf_tiny", V=320, h=64, H=1, L=2, P=48, B=3, S=7, new=5, seed=21)
    # opt-350m's shape of model: LayerNorm after the residual adds, no final LayerNorm, project_in / project_out
    gen_layer_case("layer_postln", B=2, S=6, h=128, H=2, new=2, seed=14, pre_ln=False)
    gen_hf_model_case("model_hf_postln_tiny", V=320, h=128, H=2, L=2, P=48, B=3, S=7, new=5, seed=22, word_dim=64, pre_ln=False)


if __name__ == "__main__":
    sys.exit(main())
