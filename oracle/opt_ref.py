"""CPU/GPU-agnostic ORACLE for LIA's OPT decoder-layer hot path.

TEST INFRASTRUCTURE, NOT PRODUCT.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` leg may import this
module; the product package (``isca-2025-lia_b200``) never does.

It is an op-for-op restatement, in stock PyTorch, of the reference's GPU branch
(``policy == 3``) and, separately, of its full-CPU branch (``policy == 1``).
Each function cites the reference lines it follows.  Abbreviations:

  D  = intel_extension_for_pytorch/transformers/models/reference/modules/decoder.py
  A  = intel_extension_for_pytorch/transformers/models/reference/modules/attentions.py
  M  = lia/modeling_opt.py
  RM = intel_extension_for_pytorch/transformers/models/reference/models.py
  GS = intel_extension_for_pytorch/transformers/generation/greedy_search.py
  GU = lia/generation_utils.py

Parity pin: ``tests/test_oracle_golden.py`` checks ``layer_forward`` BIT-EXACTLY
against outputs of the reference's own ``OPTDecoderLayer_forward`` /
``_OPTAttention_forward`` executed in the build container
(``oracle/gen_golden.py`` -> ``tests/golden/layer_*.npz``) and the whole-model
functions against stock ``transformers.OPTForCausalLM`` (``model_hf_tiny.npz``);
``positions_from_mask`` / ``prepare_attention_mask`` against the reference's own
``OPTLearnedPositionalEmbedding.forward`` and ``_prepare_attention_mask_for_generation``
run on padded prompts (``positions_padded.npz``); ``shard_layer`` against the reference's own tensor-parallel
sharder (``tp_shard.npz``); the post-LN branch against the reference's own layer code with
``do_layer_norm_before=False`` (``layer_postln.npz``) and stock transformers (``model_hf_postln_tiny.npz``).
The reference holds no golden vectors of its own for this path.

Weight dictionaries use the keys
  ln1_w ln1_b q_w q_b k_w k_b v_w v_b o_w o_b ln2_w ln2_b fc1_w fc1_b fc2_w fc2_b
(index order of the reference's 16-entry ``gpu_layer`` list, M:272-293) and
model dictionaries add ``embed_tokens`` [V,h], ``embed_positions`` [P+2,h],
``final_ln_w``, ``final_ln_b`` and ``layers`` (list of layer dicts).
"""
import math
import time

import torch
from torch.nn import functional as F

LN_EPS = 1e-5          # nn.LayerNorm default, M:782-787 (elementwise_affine from config, eps not overridden)
MASK_VALUE = -3.4028e+38   # A:447


# --------------------------------------------------------------------------- layer (policy 3)

def _linear(x, w, b):
    """``torch.matmul(x, w.t()) + b`` -- two roundings in bf16 (A:393-394, 418; D:81, 88)."""
    return torch.matmul(x, w.t()) + b


def attention_forward(x, w, H, kcache, vcache, cur_len):
    """A:312-557, GPU branch with policy 3.

    x [B,S,h] (already layer-normed); kcache/vcache [Tmax,B,H,d] time-major,
    ``cur_len`` tokens valid.  Writes the new K/V rows into the caches in place
    (A:462-491) and returns the context [B,S,h].
    """
    B, S, h = x.shape
    d = h // H
    scaling = d ** -0.5                                                  # M:413
    x2 = x.view(B * S, h)                                                # A:379
    key = _linear(x2, w["k_w"], w["k_b"]).view(B, S, H, d).contiguous()  # A:393
    value = _linear(x2, w["v_w"], w["v_b"]).view(B, S, H, d).contiguous()  # A:394
    if S == 1:                                                           # A:397-399
        key = torch.cat([kcache[:cur_len].permute(1, 0, 2, 3), key], dim=1).contiguous()
        value = torch.cat([vcache[:cur_len].permute(1, 0, 2, 3), value], dim=1).contiguous()
    query = _linear(x2, w["q_w"], w["q_b"]).view(B, S, H, d).contiguous()  # A:418
    T = key.size(1)
    mask = None
    if True:                                                             # A:444-449 (mask is rebuilt, never the HF one)
        mask = torch.triu(torch.ones(S, T, device=x.device), diagonal=1) * MASK_VALUE
        mask = mask.unsqueeze(0).unsqueeze(0).expand(B, 1, S, T)
    query = query * scaling                                              # A:456
    key_buff = key.permute(1, 0, 2, 3).to(torch.bfloat16) if x.dtype == torch.bfloat16 else key.permute(1, 0, 2, 3)
    value_buff = value.permute(1, 0, 2, 3).to(torch.bfloat16) if x.dtype == torch.bfloat16 else value.permute(1, 0, 2, 3)
    if S != 1:                                                           # A:462-476
        kcache[:S] = key_buff
        vcache[:S] = value_buff
    else:                                                                # A:478-491
        kcache[cur_len:cur_len + 1].copy_(key_buff[cur_len:cur_len + 1])
        vcache[cur_len:cur_len + 1].copy_(value_buff[cur_len:cur_len + 1])
    q = query.transpose(1, 2).contiguous().view(B * H, -1, d)            # A:493-496
    k = key.transpose(1, 2).contiguous().view(B * H, -1, d)
    v = value.transpose(1, 2).contiguous().view(B * H, -1, d)
    attn = torch.bmm(q, k.transpose(1, 2))                               # A:499
    if S != 1:                                                           # A:500-509 (no mask at all in decode)
        attn = attn.view(B, H, S, T) + mask
        attn = torch.max(attn, torch.tensor(torch.finfo(attn.dtype).min, device=attn.device))
    attn = attn.view(B * H, S, T)
    attn = torch.softmax(attn, dim=-1, dtype=x.dtype)                    # A:512 (casts scores to bf16 first)
    ctx = torch.bmm(attn, v)                                             # A:529
    ctx = ctx.view(B, H, S, d).transpose(1, 2).reshape(B, S, h)          # A:544-550
    return ctx


def layer_forward(x, w, H, kcache, vcache, cur_len, pre_ln=True):
    """D:172-335 with policy 3, non-distributed.  ``pre_ln`` is ``do_layer_norm_before``: True for every OPT
    size except opt-350m, whose LayerNorms follow the residual adds instead (D:250-259, D:320-321).

    Returns the new hidden state [B,S,h]; caches are updated in place.
    """
    h = x.shape[-1]
    residual = x                                                          # D:195
    y = x
    if pre_ln:
        y = F.layer_norm(y, (h,), w["ln1_w"], w["ln1_b"], LN_EPS)         # D:198-204 -> D:107-112
    y = attention_forward(y, w, H, kcache, vcache, cur_len)              # D:210-220
    y = _linear(y, w["o_w"], w["o_b"]).contiguous()                      # D:228 -> D:86-90
    y = residual + y                                                      # D:229
    if not pre_ln:
        y = F.layer_norm(y, (h,), w["ln1_w"], w["ln1_b"], LN_EPS)         # D:250-256 -> D:107-112
    residual = y                                                          # D:264
    z = y
    if pre_ln:
        z = F.layer_norm(z, (h,), w["ln2_w"], w["ln2_b"], LN_EPS)         # D:266-272 -> D:114-119
    z = F.relu(_linear(z, w["fc1_w"], w["fc1_b"]).contiguous())          # D:285 -> D:100-105
    z = _linear(z, w["fc2_w"], w["fc2_b"]).contiguous()                  # D:309
    out = (residual + z).view(x.shape)                                   # D:310
    if not pre_ln:
        out = F.layer_norm(out, (h,), w["ln2_w"], w["ln2_b"], LN_EPS)     # D:320-321 (the nn.LayerNorm module itself)
    return out


# --------------------------------------------------------------------------- whole model

def positions_from_mask(attention_mask, past_len):
    """M:368-378 (OPTLearnedPositionalEmbedding.forward), offset 2 included."""
    am = attention_mask.long()
    pos = (torch.cumsum(am, dim=1).type_as(am) * am).long() - 1
    return pos[:, past_len:] + 2


def embed(model, input_ids, attention_mask, past_len):
    """M:1107-1142: token embedding (+ ``project_in`` when word_embed_proj_dim != hidden_size, M:1139-1140:
    opt-350m only) + learned positional embedding."""
    tok = F.embedding(input_ids, model["embed_tokens"])
    pos = F.embedding(positions_from_mask(attention_mask, past_len), model["embed_positions"])
    if model.get("project_in") is not None:
        tok = torch.matmul(tok, model["project_in"].t())                  # nn.Linear(bias=False), M:994
    return tok + pos


def new_cache(model, B, Tmax, device=None, dtype=None):
    """Per-layer time-major KV cache [(S+new), B, H, d] (A:471-472, M:1277-1278)."""
    e = model["embed_positions"]          # [P+2, hidden]; embed_tokens is [V, word_embed_proj_dim]
    h, H = e.shape[1], model["H"]
    device = device or e.device
    dtype = dtype or e.dtype
    return [(torch.zeros(Tmax, B, H, h // H, dtype=dtype, device=device),
             torch.zeros(Tmax, B, H, h // H, dtype=dtype, device=device)) for _ in model["layers"]]


def decoder_forward(model, input_ids, attention_mask, cache, past_len, collect=None):
    """M:1021-1586 reduced to its intended math for fully resident layers
    (per-layer dispatch M:1246-1260; final LN M:1563-1564).  ``collect``, if a
    list, receives every layer's output hidden state."""
    x = embed(model, input_ids, attention_mask, past_len)
    pre_ln = model.get("pre_ln", True)
    for li, w in enumerate(model["layers"]):
        x = layer_forward(x, w, model["H"], cache[li][0], cache[li][1], past_len, pre_ln)
        if collect is not None:
            collect.append(x)
    h = x.shape[-1]
    if model.get("final_ln_w") is not None:                               # M:1001-1006: absent when not pre-LN
        x = F.layer_norm(x, (h,), model["final_ln_w"], model["final_ln_b"], LN_EPS)   # M:1563-1564
    if model.get("project_out") is not None:
        x = torch.matmul(x, model["project_out"].t())                     # M:1566-1567
    return x


def lm_logits(model, hidden):
    """RM:423-431: last position only, tied lm_head (M:1660), no bias."""
    return torch.matmul(hidden[:, -1:, :], model["embed_tokens"].t()).contiguous()


def prepare_attention_mask(input_ids, pad_token_id=1, eos_token_id=2):
    """GU:469-485 (_prepare_attention_mask_for_generation): when the caller passes no mask, pad ids in
    the prompt define it, provided pad != eos; otherwise all ones."""
    if pad_token_id is not None and bool((input_ids == pad_token_id).any()) and pad_token_id != eos_token_id:
        return input_ids.ne(pad_token_id).long()
    return torch.ones(input_ids.shape[:2], dtype=torch.long, device=input_ids.device)


def greedy_generate(model, input_ids, max_new_tokens, eos_token_id=2, collect_logits=None, attention_mask=None,
                    pad_token_id=1):
    """GS:144-429 with the benchmark's kwargs (RG:179-182): greedy, min_new_tokens ==
    max_new_tokens so eos is suppressed on every step (GU:872-880), stop on length
    only (GS:425).  Returns ids [B, S+new].  The mask (given, or derived as GU:469-485 does) only
    moves the learned positions (M:368-378): the GPU branch's attention ignores padding (A:446-449, A:500)."""
    B, S = input_ids.shape
    cache = new_cache(model, B, S + max_new_tokens)
    ids = input_ids
    mask = (attention_mask.long() if attention_mask is not None
            else prepare_attention_mask(input_ids, pad_token_id, eos_token_id))
    past = 0
    cur = input_ids
    for _ in range(max_new_tokens):
        hid = decoder_forward(model, cur, mask, cache, past)
        logits = lm_logits(model, hid)[:, -1, :].float()                  # GS:367
        if collect_logits is not None:
            collect_logits.append(logits.clone())
        logits[:, eos_token_id] = -float("inf")                           # GU:872-880
        nxt = torch.argmax(logits, dim=-1)                                # GS:395
        ids = torch.cat([ids, nxt[:, None]], dim=-1)                      # GS:408
        past += cur.shape[1]
        cur = nxt[:, None]
        mask = torch.cat([mask, mask.new_ones(B, 1)], dim=-1)             # GS:411
    return ids


# --------------------------------------------------------------------------- TP restatement

def shard_layer(w, H, rank, world):
    """Sharding rule of intel_extension_for_pytorch/transformers/tensor_parallel.py:30-141:
    q/k/v column split by heads, out_proj/fc2 row split, fc1 column split; the
    row-parallel bias is divided by world (tensor_parallel.py:134; D:21)."""
    h = w["q_w"].shape[0]
    f = w["fc1_w"].shape[0]
    hs, fs = h // world, f // world
    s = {k: w[k] for k in ("ln1_w", "ln1_b", "ln2_w", "ln2_b")}
    for n in ("q", "k", "v"):
        s[n + "_w"] = w[n + "_w"][rank * hs:(rank + 1) * hs].contiguous()
        s[n + "_b"] = w[n + "_b"][rank * hs:(rank + 1) * hs].contiguous()
    s["o_w"] = w["o_w"][:, rank * hs:(rank + 1) * hs].contiguous()
    s["o_b"] = w["o_b"]
    s["fc1_w"] = w["fc1_w"][rank * fs:(rank + 1) * fs].contiguous()
    s["fc1_b"] = w["fc1_b"][rank * fs:(rank + 1) * fs].contiguous()
    s["fc2_w"] = w["fc2_w"][:, rank * fs:(rank + 1) * fs].contiguous()
    s["fc2_b"] = w["fc2_b"]
    return s


# --------------------------------------------------------------------------- policy 1 (full CPU) baseline

class CpuPolicy1Runner:
    """Restatement of the reference's full-CPU policy (prefill-policy 1, decoding-policy 1;
    SURVEY.md 3.5): IPEX cannot be built here, so its operators are replaced by the
    stock PyTorch CPU ops that implement the same algorithm --
      tpp_linear_bias / tpp_linear_relu / tpp_linear_add  -> F.linear (oneDNN, AMX bf16)
        (nn/utils/_weight_prepack.py:249-281, cpu/fusions/linear_fusion.py:46-118)
      first-token attention -> F.scaled_dot_product_attention (IPEX registers its kernel as
        the CPU flash-attention override, csrc/cpu/aten/FlashAttention.cpp:30-34)
      next-token attention over the time-major cache -> q.K / fp32 softmax / P.V
        (csrc/cpu/aten/kernels/MaskedMultiHeadAttentionKrnl.cpp:513-842)
    Used only as the *reported* CPU baseline; ``kind`` = "port".
    """

    def __init__(self, h, H, f, n_layers, B, Tmax, seed=0, dtype=torch.bfloat16):
        g = torch.Generator().manual_seed(seed)
        self.h, self.H, self.d, self.B = h, H, h // H, B
        self.layers = []
        for _ in range(n_layers):
            w = {}
            for n, shape in (("q", (h, h)), ("k", (h, h)), ("v", (h, h)), ("o", (h, h)), ("fc1", (f, h)), ("fc2", (h, f))):
                w[n + "_w"] = (torch.randn(shape, generator=g) * 0.02).to(dtype)
                w[n + "_b"] = torch.zeros(shape[0], dtype=dtype)
            for n in ("ln1", "ln2"):
                w[n + "_w"] = torch.ones(h, dtype=dtype); w[n + "_b"] = torch.zeros(h, dtype=dtype)
            self.layers.append(w)
        self.cache = [(torch.zeros(Tmax, B, H, self.d, dtype=dtype), torch.zeros(Tmax, B, H, self.d, dtype=dtype))
                      for _ in range(n_layers)]

    def layer(self, x, w, kc, vc, cur_len):
        B, S, h = x.shape
        H, d = self.H, self.d
        res = x
        y = F.layer_norm(x, (h,), w["ln1_w"], w["ln1_b"], LN_EPS)
        q = F.linear(y, w["q_w"], w["q_b"]).view(B, S, H, d)
        k = F.linear(y, w["k_w"], w["k_b"]).view(B, S, H, d)
        v = F.linear(y, w["v_w"], w["v_b"]).view(B, S, H, d)
        kc[cur_len:cur_len + S] = k.permute(1, 0, 2, 3)
        vc[cur_len:cur_len + S] = v.permute(1, 0, 2, 3)
        if S != 1:
            ctx = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=True)
            ctx = ctx.transpose(1, 2).reshape(B, S, h)
        else:
            T = cur_len + 1
            kk = kc[:T].permute(1, 2, 0, 3)                    # [B,H,T,d] view of the time-major cache
            vv = vc[:T].permute(1, 2, 0, 3)
            sc = torch.matmul(q.transpose(1, 2), kk.transpose(-1, -2)).float() * (d ** -0.5)
            p = torch.softmax(sc, dim=-1).to(x.dtype)
            ctx = torch.matmul(p, vv).transpose(1, 2).reshape(B, S, h)
        y = F.linear(ctx, w["o_w"], w["o_b"]) + res
        res = y
        z = F.relu(F.linear(F.layer_norm(y, (h,), w["ln2_w"], w["ln2_b"], LN_EPS), w["fc1_w"], w["fc1_b"]))
        return F.linear(z, w["fc2_w"], w["fc2_b"]) + res

    def run(self, S, new):
        """One generate-shaped pass (prefill S tokens then ``new``-1 decode steps + the step
        that produced the first token) over the sample layers; returns seconds."""
        g = torch.Generator().manual_seed(1)
        x = torch.randn(self.B, S, self.h, generator=g).to(torch.bfloat16)
        t0 = time.perf_counter()
        with torch.inference_mode():
            cur = 0
            y = x
            for (w, (kc, vc)) in zip(self.layers, self.cache):
                y = self.layer(y, w, kc, vc, cur)
            cur = S
            xd = y[:, -1:, :].contiguous()
            for _ in range(new - 1):
                y = xd
                for (w, (kc, vc)) in zip(self.layers, self.cache):
                    y = self.layer(y, w, kc, vc, cur)
                cur += 1
        return time.perf_counter() - t0


def model_from_hf_state_dict(sd, H):
    """Build an oracle model dict from HF OPT state-dict names."""
    L = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("model.decoder.layers."))
    layers = []
    for i in range(L):
        p = f"model.decoder.layers.{i}."
        layers.append({
            "ln1_w": sd[p + "self_attn_layer_norm.weight"], "ln1_b": sd[p + "self_attn_layer_norm.bias"],
            "q_w": sd[p + "self_attn.q_proj.weight"], "q_b": sd[p + "self_attn.q_proj.bias"],
            "k_w": sd[p + "self_attn.k_proj.weight"], "k_b": sd[p + "self_attn.k_proj.bias"],
            "v_w": sd[p + "self_attn.v_proj.weight"], "v_b": sd[p + "self_attn.v_proj.bias"],
            "o_w": sd[p + "self_attn.out_proj.weight"], "o_b": sd[p + "self_attn.out_proj.bias"],
            "ln2_w": sd[p + "final_layer_norm.weight"], "ln2_b": sd[p + "final_layer_norm.bias"],
            "fc1_w": sd[p + "fc1.weight"], "fc1_b": sd[p + "fc1.bias"],
            "fc2_w": sd[p + "fc2.weight"], "fc2_b": sd[p + "fc2.bias"],
        })
    return {"H": H, "layers": layers,
            "embed_tokens": sd["model.decoder.embed_tokens.weight"],
            "embed_positions": sd["model.decoder.embed_positions.weight"],
            # opt-350m: no final LayerNorm (M:1001-1006), LayerNorm after the residual adds, and a projection either side
            "pre_ln": "model.decoder.final_layer_norm.weight" in sd,
            "final_ln_w": sd.get("model.decoder.final_layer_norm.weight"),
            "final_ln_b": sd.get("model.decoder.final_layer_norm.bias"),
            "project_in": sd.get("model.decoder.project_in.weight"),
            "project_out": sd.get("model.decoder.project_out.weight")}


def model_to(model, device=None, dtype=None):
    def cv(t):
        return t.to(device=device, dtype=dtype) if torch.is_tensor(t) else t
    out = {k: cv(v) for k, v in model.items() if k != "layers"}
    out["layers"] = [{k: cv(v) for k, v in w.items()} for w in model["layers"]]
    return out
