"""Per-layer weight slabs and deterministic weight generators.

A decoder layer's 16 tensors (the reference's ``gpu_layer`` list, lia/modeling_opt.py:272-293)
are packed into ONE contiguous bf16 slab with q/k/v fused row-wise into a [3*hq, h] matrix, so a
streamed layer is a single cudaMemcpyAsync (the reference issues 16 copy_() calls per layer,
lia/modeling_opt.py:295-318) and the QKV projection is a single GEMM.  Weights stay plain
row-major [N, K] -- already the K-major operand tcgen05 wants; the reference's IPEX 5-D
blocked layout and its per-call un-blocking do not exist here.

Tensor-parallel sharding follows intel_extension_for_pytorch/transformers/tensor_parallel.py:30-141:
q/k/v and fc1 are split by output rows (whole heads), out_proj and fc2 by input columns, and
the row-parallel biases are divided by the world size (tensor_parallel.py:134, decoder.py:21).
"""
from collections import OrderedDict

import torch

BF16 = torch.bfloat16
LAYER_KEYS = ("ln1_w", "ln1_b", "q_w", "q_b", "k_w", "k_b", "v_w", "v_b", "o_w", "o_b",
              "ln2_w", "ln2_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b")   # gpu_layer index order


def padded_head_dim(d):
    """The attention kernels are built for head_dim 64 and 128; any other head_dim that is a multiple of 8 (opt-2.7b: 80)
    runs on them ZERO-PADDED: q/k/v get zero rows and out_proj zero columns per head, so the padded lanes of q.k, of the
    cached K/V and of the context are exact zeros and every real value is what the unpadded math gives."""
    if d < 8 or d % 8 or d > 128:
        raise NotImplementedError(f"head_dim {d} unsupported: the kernels handle multiples of 8 up to 128 (padded to 64 / 128)")
    return 64 if d <= 64 else 128


class LayerLayout:
    """Element offsets of one (possibly TP-sharded) layer inside its slab.  ``heads`` (total attention heads) enables
    head padding for head_dims other than 64 / 128; ``hq`` is then the PADDED local attention width."""

    def __init__(self, h, f, tp_world=1, heads=None):
        assert h % tp_world == 0 and f % tp_world == 0
        self.h, self.f, self.tp = h, f, tp_world
        self.hq, self.fq = h // tp_world, f // tp_world
        self.heads_local = self.d = self.dp = None
        if heads is not None:
            assert h % heads == 0
            self.d = h // heads
            self.dp = padded_head_dim(self.d)
            if heads % tp_world == 0:
                self.heads_local = heads // tp_world
                self.hq = self.heads_local * self.dp           # == h // tp_world unless the heads are padded
            elif self.dp != self.d:
                raise ValueError(f"{heads} heads of padded head_dim {self.d} do not split over {tp_world} ranks")
        shapes = OrderedDict([
            ("ln1_w", (h,)), ("ln1_b", (h,)),
            ("qkv_w", (3 * self.hq, h)), ("qkv_b", (3 * self.hq,)),
            ("o_w", (h, self.hq)), ("o_b", (h,)),
            ("ln2_w", (h,)), ("ln2_b", (h,)),
            ("fc1_w", (self.fq, h)), ("fc1_b", (self.fq,)),
            ("fc2_w", (h, self.fq)), ("fc2_b", (h,)),
        ])
        self.shapes = shapes
        self.offsets = {}
        off = 0
        for k, shp in shapes.items():
            n = 1
            for s in shp:
                n *= s
            assert n % 8 == 0, (k, shp)           # every tensor starts 16-byte aligned
            self.offsets[k] = (off, n)
            off += n
        self.numel = off
        self.nbytes = off * 2

    def views(self, slab):
        """Name -> tensor views into a flat bf16 slab (no copies)."""
        assert slab.dtype == BF16 and slab.numel() >= self.numel
        return {k: slab[o:o + n].view(self.shapes[k]) for k, (o, n) in self.offsets.items()}


def shard_layer(w, rank, world):
    """Full layer dict -> this rank's shard (keys as in LAYER_KEYS)."""
    if world == 1:
        return w
    h = w["q_w"].shape[0]
    f = w["fc1_w"].shape[0]
    hs, fs = h // world, f // world
    s = {k: w[k] for k in ("ln1_w", "ln1_b", "ln2_w", "ln2_b")}
    for n in ("q", "k", "v"):
        s[n + "_w"] = w[n + "_w"][rank * hs:(rank + 1) * hs]
        s[n + "_b"] = w[n + "_b"][rank * hs:(rank + 1) * hs]
    s["o_w"] = w["o_w"][:, rank * hs:(rank + 1) * hs]
    s["o_b"] = (w["o_b"].float() / world).to(w["o_b"].dtype)
    s["fc1_w"] = w["fc1_w"][rank * fs:(rank + 1) * fs]
    s["fc1_b"] = w["fc1_b"][rank * fs:(rank + 1) * fs]
    s["fc2_w"] = w["fc2_w"][:, rank * fs:(rank + 1) * fs]
    s["fc2_b"] = (w["fc2_b"].float() / world).to(w["fc2_b"].dtype)
    return s


def _pad_head_rows(w, Hl, d, dp):
    """[Hl*d, ...] -> [Hl*dp, ...]: zero rows after each head's d rows."""
    out = w.new_zeros((Hl, dp) + tuple(w.shape[1:]))
    out[:, :d] = w.reshape((Hl, d) + tuple(w.shape[1:]))
    return out.reshape((Hl * dp,) + tuple(w.shape[1:]))


def _pad_head_cols(w, Hl, d, dp):
    """[N, Hl*d] -> [N, Hl*dp]: zero columns after each head's d columns."""
    out = w.new_zeros(w.shape[0], Hl, dp)
    out[:, :, :d] = w.reshape(w.shape[0], Hl, d)
    return out.reshape(w.shape[0], Hl * dp)


def fuse_layer(s, layout):
    """This rank's (already sharded) 16-tensor layer dict -> the fused tensors the kernels consume: q|k|v stacked row-wise
    (one QKV GEMM) and, where ``layout`` pads the heads, zero rows / columns per head (see padded_head_dim)."""
    q = [s["q_w"], s["k_w"], s["v_w"]]
    b = [s["q_b"], s["k_b"], s["v_b"]]
    o = s["o_w"]
    if layout.dp is not None and layout.dp != layout.d:
        Hl, d, dp = layout.heads_local, layout.d, layout.dp
        q = [_pad_head_rows(t, Hl, d, dp) for t in q]
        b = [_pad_head_rows(t, Hl, d, dp) for t in b]
        o = _pad_head_cols(o, Hl, d, dp)
    out = {k: s[k] for k in ("ln1_w", "ln1_b", "o_b", "ln2_w", "ln2_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b")}
    out["qkv_w"], out["qkv_b"], out["o_w"] = torch.cat(q, dim=0), torch.cat(b, dim=0), o
    return out


def pack_layer(w, layout, rank=0, out=None):
    """Pack a FULL layer dict into this rank's slab (flat bf16 tensor on w's device or ``out``)."""
    s = fuse_layer(shard_layer(w, rank, layout.tp), layout)
    dev = s["qkv_w"].device
    slab = out if out is not None else torch.empty(layout.numel, dtype=BF16, device=dev)
    v = layout.views(slab)
    for k in v:
        v[k].copy_(s[k])
    return slab


def random_layer(h, f, seed, device="cpu", kind="normal", std=0.02, bias_std=0.0, ln_std=0.0):
    """Deterministic full (unsharded) layer.

    kind="normal": lia/modeling_opt.py:895-904 (_init_weights: normal(0, std) weights, zero
    biases, LayerNorm w=1 b=0); ``bias_std``/``ln_std`` > 0 perturb biases / LN affine so tests
    exercise those terms.  kind="dummy": utils/opt-weight-gen.py:61-62 -- EVERY parameter
    ~ U[0,1) in bf16 (activations overflow quickly; throughput-only).
    """
    g = torch.Generator(device=device).manual_seed(seed)
    w = {}

    def normal(shape, s, mean=0.0):
        return (torch.randn(shape, generator=g, device=device, dtype=torch.float32) * s + mean).to(BF16)

    def uniform(shape):
        return torch.rand(shape, generator=g, device=device, dtype=torch.float32).to(BF16)

    mats = (("q", (h, h)), ("k", (h, h)), ("v", (h, h)), ("o", (h, h)), ("fc1", (f, h)), ("fc2", (h, f)))
    if kind == "dummy":
        for n in ("ln1", "ln2"):
            w[n + "_w"], w[n + "_b"] = uniform((h,)), uniform((h,))
        for n, shp in mats:
            w[n + "_w"], w[n + "_b"] = uniform(shp), uniform((shp[0],))
        return w
    for n in ("ln1", "ln2"):
        w[n + "_w"] = normal((h,), ln_std, 1.0) if ln_std > 0 else torch.ones(h, dtype=BF16, device=device)
        w[n + "_b"] = normal((h,), ln_std) if ln_std > 0 else torch.zeros(h, dtype=BF16, device=device)
    for n, shp in mats:
        w[n + "_w"] = normal(shp, std)
        w[n + "_b"] = normal((shp[0],), bias_std) if bias_std > 0 else torch.zeros(shp[0], dtype=BF16, device=device)
    return w


def random_embeddings(vocab, h, max_pos, seed, device="cpu", kind="normal", std=0.02, ln_std=0.0, pad_id=1, embed_dim=None,
                      final_ln=True):
    """embed_tokens [V,e] (padding row zeroed, lia/modeling_opt.py:900-904), embed_positions
    [max_pos+2, h] (offset 2, :365-366), final LayerNorm (pre-LN models only, :1001-1006) and, when the token
    table is narrower than the model (e != h: opt-350m, :988-996), bias-free project_in [h,e] / project_out [e,h].
    The generator is consumed in the same order as before for e == h, so existing seeds give the same weights."""
    g = torch.Generator(device=device).manual_seed(seed)
    e_dim = embed_dim or h
    if kind == "dummy":
        r = lambda *s: torch.rand(*s, generator=g, device=device, dtype=torch.float32).to(BF16)  # noqa: E731
        e = {"embed_tokens": r(vocab, e_dim), "embed_positions": r(max_pos + 2, h)}
        if final_ln:
            e["final_ln_w"], e["final_ln_b"] = r(h), r(h)
        if e_dim != h:
            e["project_in"], e["project_out"] = r(h, e_dim), r(e_dim, h)
        return e
    n = lambda *s: (torch.randn(*s, generator=g, device=device, dtype=torch.float32) * std).to(BF16)  # noqa: E731
    e = {"embed_tokens": n(vocab, e_dim), "embed_positions": n(max_pos + 2, h)}
    e["embed_tokens"][pad_id].zero_()
    if final_ln:
        if ln_std > 0:
            e["final_ln_w"] = (1 + torch.randn(h, generator=g, device=device) * ln_std).to(BF16)
            e["final_ln_b"] = (torch.randn(h, generator=g, device=device) * ln_std).to(BF16)
        else:
            e["final_ln_w"] = torch.ones(h, dtype=BF16, device=device)
            e["final_ln_b"] = torch.zeros(h, dtype=BF16, device=device)
    if e_dim != h:
        e["project_in"], e["project_out"] = n(h, e_dim), n(e_dim, h)
    return e


def layer_from_hf_state_dict(sd, i):
    """HF OPT state-dict names -> layer dict."""
    p = f"model.decoder.layers.{i}."
    return {
        "ln1_w": sd[p + "self_attn_layer_norm.weight"], "ln1_b": sd[p + "self_attn_layer_norm.bias"],
        "q_w": sd[p + "self_attn.q_proj.weight"], "q_b": sd[p + "self_attn.q_proj.bias"],
        "k_w": sd[p + "self_attn.k_proj.weight"], "k_b": sd[p + "self_attn.k_proj.bias"],
        "v_w": sd[p + "self_attn.v_proj.weight"], "v_b": sd[p + "self_attn.v_proj.bias"],
        "o_w": sd[p + "self_attn.out_proj.weight"], "o_b": sd[p + "self_attn.out_proj.bias"],
        "ln2_w": sd[p + "final_layer_norm.weight"], "ln2_b": sd[p + "final_layer_norm.bias"],
        "fc1_w": sd[p + "fc1.weight"], "fc1_b": sd[p + "fc1.bias"],
        "fc2_w": sd[p + "fc2.weight"], "fc2_b": sd[p + "fc2.bias"],
    }
