"""OPT model surface of LIA, re-built on the B200 kernels.

Keeps the reference's faces:
  * ``OPTForCausalLM.forward(input_ids, attention_mask, past_key_values, ..., prefill_policy,
    decoding_policy, no_overlap, pin_weight, gpu_percentage, num_minibatch, enable_cxl,
    max_new_tokens) -> (logits [B,1,V], past_key_values)``
    (intel_extension_for_pytorch/transformers/models/reference/models.py:371-445)
  * ``OPTDecoder.forward`` with the same kwargs (lia/modeling_opt.py:1021-1586)
  * ``OPTDecoderLayer.forward(hidden_states, attention_mask, layer_head_mask, past_key_value,
    output_attentions, use_cache, gpu_layer, policy, max_new_tokens)``
    (.../reference/modules/decoder.py:172-183, 323-335)
  * ``generate(input_ids, max_new_tokens=, min_new_tokens=, do_sample=False, num_beams=1,
    <LIA kwargs>)`` (single_instance/run_generation.py:179-182, 319)
  * ``past_key_values[l] = (marker, K, V, beam_idx)`` with ``marker.shape[2] == tokens cached``
    and K/V time-major ``[S+new, B, H, d]`` (attentions.py:462-491, greedy_search.py:272-282).

What differs by design: every layer op is one call into libliab200.so (no torch math on the
path), the whole stack -- embeddings, layers, final LayerNorm, lm_head, argmax -- stays on
the GPU, buffers/streams are allocated once (the reference re-creates them per forward,
lia/modeling_opt.py:1180-1212), non-resident layers are streamed from pinned host memory
instead of being computed on the CPU, and decode steps are replayed from CUDA graphs.

Policy mapping: 0/2/3/4 -> all-GPU compute; 1 (full CPU, the IPEX/AMX
baseline) is not part of this build -- ``bench.py --impl reference`` times it.
"""
import math
import os
import time
from dataclasses import dataclass, field

import torch

from . import _lib, graphs, ops, program as program_mod, tp as tp_mod
from .ops import EPI_BIAS, EPI_BIAS_RELU, EPI_BIAS_RESIDUAL, EPI_QKV
from .kv_spill import KVSpill, plan_resident_layers
from .streamer import HostArena, LayerStreamer
from .weights import LAYER_KEYS, LayerLayout, fuse_layer, layer_from_hf_state_dict, pack_layer, random_embeddings, random_layer

BF16 = torch.bfloat16
LN_EPS = 1e-5


@dataclass
class OPTConfig:
    """Fields of the HF OPTConfig that the path reads (lia/modeling_opt.py:977-1013)."""
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    ffn_dim: int = 3072
    vocab_size: int = 50272
    max_position_embeddings: int = 2048
    do_layer_norm_before: bool = True  # False (opt-350m): LayerNorm AFTER each residual add, no final LayerNorm
    word_embed_proj_dim: int = 0       # 0 = hidden_size; opt-350m embeds in 512 and projects in/out (M:985-996)
    pad_token_id: int = 1
    bos_token_id: int = 2
    eos_token_id: int = 2
    init_std: float = 0.02
    token_latency: bool = False        # run_generation.py:153-154
    lm_head_generation: bool = True    # run_generation.py:155-156
    name: str = "opt"
    architectures: list = field(default_factory=lambda: ["OPTForCausalLM"])

    @property
    def head_dim(self):
        return self.hidden_size // self.num_attention_heads

    @property
    def embed_dim(self):
        return self.word_embed_proj_dim or self.hidden_size


def _cfg(name, L, h, H, f, **kw):
    return OPTConfig(hidden_size=h, num_hidden_layers=L, num_attention_heads=H, ffn_dim=f, name=name, **kw)


# SURVEY.md A.1; 66b/175b from utils/opt-weight-gen.py:83-131
OPT_CONFIGS = {
    "opt-125m": _cfg("opt-125m", 12, 768, 12, 3072),
    "opt-350m": _cfg("opt-350m", 24, 1024, 16, 4096, do_layer_norm_before=False, word_embed_proj_dim=512),
    "opt-1.3b": _cfg("opt-1.3b", 24, 2048, 32, 8192),
    "opt-2.7b": _cfg("opt-2.7b", 32, 2560, 32, 10240),           # head_dim 80: runs zero-padded to 128 (weights.padded_head_dim)
    "opt-6.7b": _cfg("opt-6.7b", 32, 4096, 32, 16384),
    "opt-13b": _cfg("opt-13b", 40, 5120, 40, 20480),
    "opt-30b": _cfg("opt-30b", 48, 7168, 56, 28672),
    "opt-66b": _cfg("opt-66b", 64, 9216, 72, 36864),
    "opt-175b": _cfg("opt-175b", 96, 12288, 96, 49152),
}


def get_config(name_or_path):
    key = str(name_or_path).rstrip("/").split("/")[-1].lower()
    if key not in OPT_CONFIGS:
        raise KeyError(f"unknown OPT model {name_or_path!r}; known: {sorted(OPT_CONFIGS)}")
    c = OPT_CONFIGS[key]
    return OPTConfig(**{k: getattr(c, k) for k in c.__dataclass_fields__})


def _check_policy(p, what):
    if p is None:
        return 3
    if p == 1:
        raise NotImplementedError(
            f"{what}=1 is the reference's full-CPU (IPEX/AMX) policy; this build has no CPU compute path. "
            "It is timed as the reported baseline by `bench.py --impl reference`.")
    if p not in (0, 2, 3, 4):
        raise ValueError(f"{what} must be one of 0,1,2,3,4 (got {p})")
    return p


class _Workspace:
    """Activation scratch for up to ``rows`` token rows -- allocated once per shape."""

    def __init__(self, cfg, layout, rows, batch, device, arena=None):
        h, hq, fq = cfg.hidden_size, layout.hq, layout.fq
        e = lambda *s: torch.empty(*s, dtype=BF16, device=device)  # noqa: E731
        self.rows = rows
        self.arena = arena          # tp.PeerArena / SymmArena: row-parallel projections run fused with their all-reduce
        self.arena_d = arena        # arena of the decode-shaped (M <= 128) exchanges, when it is a separate one
        self.ln, self.q, self.ctx, self.ffn = e(rows, h), e(rows, hq), e(rows, hq), e(rows, fq)
        self.x1 = arena.tensor("x1", (rows, h)) if arena is not None else e(rows, h)
        # decode (M <= 128) needs no remotely writable output: keep its residual stream out of the peer-mapped arena
        self.x1d = e(batch, h) if (arena is not None and os.environ.get("LIA_TP_DECODE_ARENA", "0") == "0") else self.x1
        self.tp = e(rows, h) if layout.tp > 1 and arena is None else None
        shapes = []
        e_dim = cfg.embed_dim
        for m in {rows, batch}:
            shapes += [(m, 3 * hq, h), (m, h, hq), (m, fq, h), (m, h, fq), (m, cfg.vocab_size, e_dim)]
            if e_dim != h:
                shapes += [(m, h, e_dim), (m, e_dim, h)]                   # project_in / project_out
        self.gemm = ops.GemmWorkspace(ops.GemmWorkspace.bytes_for(shapes), device)
        self.attn = ops.attn_decode_workspace(batch, cfg.num_attention_heads // layout.tp, layout.dp, device)


class _GenState:
    """Everything one (B, S, new) generation shape needs, kept across generate() calls."""

    def __init__(self, model, B, S, new, num_minibatch):
        cfg, dev = model.config, model.device
        L, d = cfg.num_hidden_layers, model.layout.dp        # cached rows are head-padded where head_dim is not 64 / 128
        Hl = cfg.num_attention_heads // model.tp_world
        self.B, self.S, self.new = B, S, new
        self.Tmax = S + new
        self.beam_idx = torch.zeros(self.Tmax, B, dtype=torch.long, device=dev)
        self.prompt = torch.zeros(B, S, dtype=torch.int64, device=dev)
        # attention mask of prompt + generated tokens; generated columns stay 1 (greedy_search.py:411 appends ones)
        self.mask = torch.ones(B, self.Tmax, dtype=torch.int64, device=dev)
        self.steps_tok = torch.zeros(max(new, 1), B, dtype=torch.int64, device=dev)
        mb = max(1, B // max(1, num_minibatch))
        rows = max(mb * S, B)
        h = cfg.hidden_size
        self.arena = self.arena_d = None
        if model.tp_world > 1 and tp_mod.fused_enabled() and dev.type == "cuda":
            # the residual stream lives in a peer-mapped arena: owners of a prefill tile write the reduced
            # result straight into every rank's copy (include/lia_b200.h, lia_gemm_allreduce_bf16)
            lay, lib = model.layout, _lib.load()
            recv = max(lib.lia_tp_recv_bytes(m, h, k, model.tp_world) for m in {mb * S, B} for k in (lay.hq, lay.fq))
            loop = os.environ.get("LIA_TP_SELF_LOOP", "0") != "0"      # timing probe: one process, peers = itself
            # symmetric memory with an NVLink-switch multicast mapping where the pod has one (in-switch reduction of the
            # prefill exchange), CUDA-IPC peer mappings otherwise
            arena_cls = tp_mod.SymmArena if (not loop and tp_mod.SymmArena.available(dev)) else tp_mod.PeerArena
            self.arena = arena_cls(model.tp_rank, model.tp_world, dev, recv,
                                   [("x", B * S * h * 2), ("xd", B * h * 2), ("x1", rows * h * 2)],
                                   exchange=(lambda mine: [0] * model.tp_world) if loop else None)
            if loop:
                self.arena.peers = [self.arena.local] * model.tp_world
            # decode-shaped exchanges (M <= 128: latency-bound LL stores into peers' receive areas) keep a CUDA-IPC arena of
            # their own: measured 0.25 ms per decode step slower at TP4 through memory bound to a multicast object
            # (profiles/README.md), while the prefill exchange gains from the in-switch reduction
            self.arena_d = self.arena
            if isinstance(self.arena, tp_mod.SymmArena) and S > 1 and os.environ.get("LIA_TP_DECODE_ARENA", "0") == "0":
                recv_d = max(lib.lia_tp_recv_bytes(m, h, k, model.tp_world) for m in {min(mb * S, 128), B} for k in (lay.hq, lay.fq))
                self.arena_d = tp_mod.PeerArena(model.tp_rank, model.tp_world, dev, recv_d, [("pad", 256)])
            self.x = self.arena.tensor("x", (B * S, h))
            self.xd = (self.arena.tensor("xd", (B, h)) if os.environ.get("LIA_TP_DECODE_ARENA", "0") != "0"
                       else torch.empty(B, h, dtype=BF16, device=dev))
        else:
            self.x = torch.empty(B * S, h, dtype=BF16, device=dev)
            self.xd = torch.empty(B, h, dtype=BF16, device=dev)
        self.xn = torch.empty(B, cfg.hidden_size, dtype=BF16, device=dev)
        self.logits = torch.empty(B, cfg.vocab_size, dtype=BF16, device=dev)
        self.ws = _Workspace(cfg, model.layout, rows, B, dev, self.arena)
        self.ws.arena_d = getattr(self, "arena_d", None)
        self.graphs = {}
        self.program = None         # program.DecodeProgram: the whole decode step as one persistent kernel (False: unsupported)
        self.calls = 0
        # KV cache last, once everything else of this shape is placed: layers whose K/V do not fit in HBM are
        # spilled to pinned host memory (kv_spill.py; lia/modeling_opt.py:326-349 keeps the whole cache there)
        per_layer = 2 * self.Tmax * B * Hl * d * 2
        n_res = model.kv_resident_layers
        if n_res is None and os.environ.get("LIA_KV_RESIDENT_LAYERS"):
            n_res = int(os.environ["LIA_KV_RESIDENT_LAYERS"])
        if n_res is None:
            if dev.type == "cuda":
                torch.cuda.empty_cache()
                free = torch.cuda.mem_get_info(dev)[0]
            else:
                free = L * per_layer + (8 << 30)
            n_res = plan_resident_layers(L, per_layer, free)
        n_res = max(0, min(L, int(n_res)))
        self.kv_resident = n_res
        self.kc = [torch.zeros(self.Tmax, B, Hl, d, dtype=BF16, device=dev) if i < n_res else None for i in range(L)]
        self.vc = [torch.zeros(self.Tmax, B, Hl, d, dtype=BF16, device=dev) if i < n_res else None for i in range(L)]
        self.spill = KVSpill(L - n_res, self.Tmax, B, Hl, d, dev) if n_res < L else None

    def past_key_values(self, T):
        """The reference's per-layer 4-tuple; a spilled layer's K/V are its pinned host tensors (where the
        reference keeps every layer's, M:1383-1388)."""
        marker = torch.empty(1, T, T, 1, dtype=torch.long, device="meta")   # only .shape is ever read (M:1111)
        ks = [k if k is not None else self.spill.host_k[i - self.kv_resident] for i, k in enumerate(self.kc)]
        vs = [v if v is not None else self.spill.host_v[i - self.kv_resident] for i, v in enumerate(self.vc)]
        return tuple((marker, k, v, self.beam_idx) for k, v in zip(ks, vs))

    def close(self):
        if self.program:
            self.program.close()
        self.program = None
        self.graphs = {}
        if getattr(self, "arena_d", None) is not None and self.arena_d is not self.arena:
            self.arena_d.close()
        self.arena_d = None
        if self.arena is not None:
            self.arena.close()
            self.arena = None
        if self.spill is not None:
            self.spill.close()
            self.spill = None


class OPTDecoderLayer:
    """decoder.py:172-335 (+ attentions.py:312-557) as seven kernel launches."""

    def __init__(self, decoder, idx):
        self.decoder, self.idx = decoder, idx
        self.do_layer_norm_before = decoder.config.do_layer_norm_before          # M:778
        self.distributed = decoder.tp_world > 1

    # ---- the reference's layer face
    def forward(self, hidden_states, attention_mask=None, layer_head_mask=None, past_key_value=None,
                output_attentions=False, use_cache=False, gpu_layer=None, policy=0, max_new_tokens=None):
        if layer_head_mask is not None or output_attentions:
            raise NotImplementedError("layer_head_mask / output_attentions are not supported on the GPU path")
        _check_policy(policy, "policy")
        dec = self.decoder
        B, S, h = hidden_states.shape
        views = self._views_from_gpu_layer(gpu_layer) if gpu_layer is not None else dec.layer_views(self.idx)
        past_len = 0 if past_key_value is None else int(past_key_value[0].shape[2])
        Hl, d = dec.config.num_attention_heads // dec.tp_world, dec.layout.dp
        if S != 1:
            if past_len != 0:
                raise NotImplementedError("multi-token forward with a non-empty KV cache is not supported")
            tmax = S + (max_new_tokens if max_new_tokens is not None else dec.config.max_position_embeddings - S)
            kc = torch.zeros(tmax, B, Hl, d, dtype=BF16, device=hidden_states.device)   # attentions.py:471-472
            vc = torch.zeros_like(kc)
            beam = torch.zeros(tmax, B, dtype=torch.long, device=hidden_states.device)
        else:
            kc, vc, beam = past_key_value[1], past_key_value[2], past_key_value[3]
        x = hidden_states.reshape(B * S, h).clone()
        ws = dec.workspace(B * S, B)
        dec.layer_rows(views, x, kc, vc, B, S, past_len, 0, ws)
        T = past_len + S
        present = (torch.empty(1, T, T, 1, dtype=torch.long, device="meta"), kc, vc, beam)
        outputs = (x.view(B, S, h),)
        if use_cache:
            outputs += (present,)
        if policy == 0:                                                   # decoder.py:331-333
            outputs += (kc[past_len:T], vc[past_len:T])
        return outputs

    __call__ = forward

    def _views_from_gpu_layer(self, gl):
        """16-entry list in the reference's order (lia/modeling_opt.py:272-293) -> fused views."""
        if len(gl) != 16:
            raise ValueError("gpu_layer must have 16 entries (ln1 w/b, q w/b, k w/b, v w/b, out w/b, ln2 w/b, fc1 w/b, fc2 w/b)")
        return fuse_layer(dict(zip(LAYER_KEYS, gl)), self.decoder.layout)


class OPTDecoder:
    """lia/modeling_opt.py:977-1586: embeddings, layer placement, streaming, layer loop, final LN."""

    def __init__(self, config, device, tp_rank=0, tp_world=1):
        self.config, self.device = config, torch.device(device)
        self.tp_rank, self.tp_world = tp_rank, tp_world
        if config.num_attention_heads % tp_world or config.ffn_dim % tp_world:
            raise ValueError("heads and ffn_dim must divide by the tensor-parallel world size")
        # head_dim 64 / 128 as they are; other multiples of 8 (opt-2.7b: 80) zero-padded per head (weights.padded_head_dim)
        self.layout = LayerLayout(config.hidden_size, config.ffn_dim, tp_world, heads=config.num_attention_heads)
        self.layers = [OPTDecoderLayer(self, i) for i in range(config.num_hidden_layers)]
        self.scaling = config.head_dim ** -0.5                      # lia/modeling_opt.py:413
        self.embed_tokens = self.embed_positions = self.final_ln_w = self.final_ln_b = None
        self.project_in = self.project_out = None      # [h, e] / [e, h], only when word_embed_proj_dim != hidden_size
        self.pre_ln = bool(config.do_layer_norm_before)
        self.n_resident = 0
        self.resident = []          # flat device slabs
        self.resident_views = []
        self.host_arena = None
        self.host_pool = 0
        self.host_slabs = []
        self.streamer = None
        self._ws = {}
        self.k = ops                # kernel layer the compute methods call: lia_b200.ops, or a program.ProgramRecorder while a
                                    # decode program is being built from the very same methods

    # ---- weights & placement (move_gpu_layer / pin_memory, lia/modeling_opt.py:167-268)
    def load_layers(self, layer_fn, gpu_percentage=100):
        """``layer_fn(i, device)`` -> full layer dict on ``device``; resident layers are packed on the
        GPU, the rest straight into the pinned host arena."""
        def fill(i, out):
            w = layer_fn(i, self.device)
            if out.device.type == "cpu":
                out.copy_(pack_layer(w, self.layout, self.tp_rank))
            else:
                pack_layer(w, self.layout, self.tp_rank, out=out)
        self._place(fill, gpu_percentage, host_fill=False)

    def load_packed(self, read_slab, gpu_percentage=100):
        """``read_slab(i, out)`` fills ``out`` -- a flat CPU bf16 tensor of ``layout.numel`` elements -- with
        this rank's already-packed slab of layer i (checkpoint.SlabCheckpoint.read_slab).  Streamed layers are
        read straight into the pinned arena; resident ones go through one pinned staging slab."""
        self._place(read_slab, gpu_percentage, host_fill=True)

    def _place(self, fill, gpu_percentage, host_fill):
        L = self.config.num_hidden_layers
        n_res = L if gpu_percentage >= 100 else int(L * gpu_percentage / 100)   # M:1182 (100 == intended "all")
        self.n_resident = n_res
        self.resident, self.resident_views, self.host_slabs = [], [], []
        if self.streamer is not None:
            self.streamer.close()
            self.streamer = None
        n_host = L - n_res
        # host RAM smaller than the streamed weights (OPT-175B: 348 GB): keep only `pool` DISTINCT streamed layers in
        # the pinned arena and let streamed layer j alias slab j % pool -- the bytes crossing PCIe per step are
        # unchanged, the results are not those of the full model (throughput runs on dummy weights only)
        pool = int(os.environ.get("LIA_HOST_LAYER_POOL", "0") or 0)
        self.host_pool = pool if 0 < pool < n_host else 0
        n_slabs = self.host_pool or n_host
        if n_host:
            self.host_arena = HostArena(self.layout.numel * n_slabs)
        stage = HostArena(self.layout.numel) if (host_fill and n_res) else None
        for i in range(L):
            if i < n_res:
                slab = torch.empty(self.layout.numel, dtype=BF16, device=self.device)
                if host_fill:
                    fill(i, stage.tensor)
                    slab.copy_(stage.tensor)                 # synchronous w.r.t. the host: staging is reusable
                else:
                    fill(i, slab)
                self.resident.append(slab)
                self.resident_views.append(self.layout.views(slab))
            else:
                j = i - n_res
                k = j % n_slabs
                dst = self.host_arena.tensor[k * self.layout.numel:(k + 1) * self.layout.numel]
                if j < n_slabs:
                    fill(i, dst)
                self.host_slabs.append(dst)
        if stage is not None:
            torch.cuda.synchronize(self.device)
            stage.close()
        if n_host:
            torch.cuda.synchronize(self.device)
            self.streamer = LayerStreamer(self.layout, self.host_slabs, self.device)

    def load_embeddings(self, e):
        dev, cfg = self.device, self.config
        put = lambda t: None if t is None else t.to(dev, BF16).contiguous()  # noqa: E731
        self.embed_tokens = put(e["embed_tokens"])
        self.embed_positions = put(e["embed_positions"])
        # M:1001-1006: the final LayerNorm exists only for pre-LN models
        self.final_ln_w, self.final_ln_b = put(e.get("final_ln_w")), put(e.get("final_ln_b"))
        if self.pre_ln != (self.final_ln_w is not None):
            raise ValueError(f"do_layer_norm_before={self.pre_ln} but the checkpoint "
                             f"{'has no' if self.pre_ln else 'has a'} final LayerNorm")
        # M:988-996: project_in / project_out (bias-free) either side of the stack when the tables are narrower
        self.project_in, self.project_out = put(e.get("project_in")), put(e.get("project_out"))
        h, ed = cfg.hidden_size, cfg.embed_dim
        if (ed != h) != (self.project_in is not None and self.project_out is not None):
            raise ValueError(f"word_embed_proj_dim {ed} vs hidden_size {h}: project_in/project_out "
                             f"{'missing' if ed != h else 'unexpected'}")
        if tuple(self.embed_tokens.shape) != (cfg.vocab_size, ed) or self.embed_positions.shape[1] != h:
            raise ValueError(f"embedding tables {tuple(self.embed_tokens.shape)} / {tuple(self.embed_positions.shape)} do not "
                             f"match vocab {cfg.vocab_size}, word_embed_proj_dim {ed}, hidden_size {h}")

    def embed_rows(self, ids, past_len, mask, out, ws):
        """hidden = project_in(embed_tokens[ids]) + embed_positions[pos]  (M:1107-1142) into ``out`` [B,S,h]."""
        if self.project_in is None:
            return self.k.embed(ids, self.embed_tokens, self.embed_positions, past_len, out=out, attention_mask=mask)
        B, S = ids.shape
        h, ed = self.config.hidden_size, self.config.embed_dim
        tok = self.k.embed(ids, self.embed_tokens, None, past_len, attention_mask=mask)      # token rows [B,S,e] (no table = +0)
        pos = self.k.embed(ids, None, self.embed_positions, past_len, attention_mask=mask)   # position rows [B,S,h]
        self.k.gemm(tok.view(B * S, ed), self.project_in, None, out=out.view(B * S, h), epilogue=EPI_BIAS_RESIDUAL,
                 residual=pos.view(B * S, h), workspace=ws.gemm)                                   # M:1139-1142
        return out

    def final_rows(self, rows, ws, out=None):
        """Final LayerNorm (pre-LN models, M:1563-1564) and project_out (M:1566-1567) over ``rows`` [M,h]."""
        if self.final_ln_w is not None:
            rows = self.k.layernorm(rows, self.final_ln_w, self.final_ln_b, LN_EPS, out=out)
        if self.project_out is not None:
            rows = self.k.gemm(rows, self.project_out, None, epilogue=EPI_BIAS, workspace=ws.gemm)
        return rows

    def check_gpu_percentage(self, gpu_percentage):
        if gpu_percentage is None:
            return
        L = self.config.num_hidden_layers
        want = L if gpu_percentage >= 100 else int(L * gpu_percentage / 100)
        if want != self.n_resident:
            raise ValueError(f"model was placed with {self.n_resident} resident layers but gpu_percentage="
                             f"{gpu_percentage} asks for {want}; reload with load_layers(..., gpu_percentage=)")

    def layer_views(self, i):
        if i < self.n_resident:
            return self.resident_views[i]
        raise RuntimeError(f"layer {i} is streamed; call it through OPTDecoder.forward")

    def workspace(self, rows, batch):
        key = (rows, batch)
        if key not in self._ws:
            self._ws[key] = _Workspace(self.config, self.layout, rows, batch, self.device)
        return self._ws[key]

    # ---- one layer over a block of token rows (rows are [b-major, S] and updated in place)
    def _row_parallel(self, a, w, b, residual, out, ws, big):
        """out = residual + (a . w^T + b), summed over the tensor-parallel ranks (out_proj D:222-247, fc2 D:302-317)."""
        arena = ws.arena
        if self.tp_world == 1:
            self.k.gemm(a, w, b, out=out, epilogue=EPI_BIAS_RESIDUAL, residual=residual, workspace=ws.gemm)   # D:228-229, 309-310
        elif arena is not None:             # D:60-68 + 247/317 as one kernel over NVLink peer memory
            # prefill tiles are finished by their owner rank, which writes `out` remotely: it must live in the arena
            args = arena.args(out) if big else (ws.arena_d or arena).args(None)
            self.k.gemm_allreduce(a, w, b, residual, out, args, workspace=ws.gemm)
        else:
            part = ws.tp[:a.shape[0]]
            self.k.gemm(a, w, b, out=part, epilogue=EPI_BIAS, workspace=ws.gemm)                              # D:60-68
            tp_mod.all_reduce(part)
            self.k.residual_add(part, residual, out=out)                                                      # D:247, 317

    def layer_rows(self, v, rows, kc, vc, nb, S, pos0, b0, ws):
        M = nb * S
        ln, q, ctx, ffn = ws.ln[:M], ws.q[:M], ws.ctx[:M], ws.ffn[:M]
        big = M > 128
        # decode (S == 1) keeps its residual stream out of the peer-mapped arena; a short PREFILL (M <= 128, S > 1) has up to
        # 128 rows and uses the prefill buffer (x1d holds `batch` rows only)
        x1 = ws.x1d[:M] if S == 1 else ws.x1[:M]
        pre = self.pre_ln
        if pre:
            self.k.layernorm(rows, v["ln1_w"], v["ln1_b"], LN_EPS, out=ln)                               # decoder.py:198-204
        self.k.gemm(ln if pre else rows, v["qkv_w"], v["qkv_b"], epilogue=EPI_QKV,                       # attentions.py:376-491
                 qkv=self.k.qkv_args(q, kc, vc, S, pos0, b0, self.scaling), workspace=ws.gemm)
        if S != 1:
            self.k.attn_prefill(q, kc, vc, nb, S, b0, out=ctx)                                           # attentions.py:493-536
        else:
            # a tensor-parallel rank keeps H / world heads: ask for more key ranges per (b, h) than the unsharded default
            self.k.attn_decode(q, kc, vc, nb, pos0 + 1, b0, out=ctx, splits=-6 if self.layout.tp > 1 else 0, workspace=ws.attn)
        self._row_parallel(ctx, v["o_w"], v["o_b"], rows, x1, ws, big)                                # decoder.py:222-247
        if pre:
            self.k.layernorm(x1, v["ln2_w"], v["ln2_b"], LN_EPS, out=ln)                                 # decoder.py:266-272
            self.k.gemm(ln, v["fc1_w"], v["fc1_b"], out=ffn, epilogue=EPI_BIAS_RELU, workspace=ws.gemm)  # decoder.py:285
            self._row_parallel(ffn, v["fc2_w"], v["fc2_b"], x1, rows, ws, big)                        # decoder.py:302-317
        else:
            # opt-350m: LayerNorm follows each residual add; its output is both the MLP input and the next residual
            self.k.layernorm(x1, v["ln1_w"], v["ln1_b"], LN_EPS, out=ln)                                 # decoder.py:250-256
            self.k.gemm(ln, v["fc1_w"], v["fc1_b"], out=ffn, epilogue=EPI_BIAS_RELU, workspace=ws.gemm)  # decoder.py:285
            self._row_parallel(ffn, v["fc2_w"], v["fc2_b"], ln, x1, ws, big)                          # decoder.py:302-317
            self.k.layernorm(x1, v["ln2_w"], v["ln2_b"], LN_EPS, out=rows)                               # decoder.py:320-321

    def set_overlap(self, overlap, spill=None):
        """``--no-overlap`` (M:1173): transfers of streamed layers / spilled K/V are not fetched ahead but serialised
        with the compute that uses them -- the reference's ablation of its prefetch pipeline.  Results are unchanged."""
        if self.streamer is not None:
            self.streamer.overlap = bool(overlap)
        if spill is not None:
            spill.overlap = bool(overlap)

    def run_layers(self, x, kcs, vcs, B, S, pos0, num_minibatch, ws, spill=None):
        """Layer-major, minibatch-minor loop (lia/modeling_opt.py:1222, 1284) with double-buffered
        weight streaming for non-resident layers and, where ``kcs[li] is None``, double-buffered K/V
        of that layer from the pinned host spill (``spill``: kv_spill.KVSpill holding the last layers)."""
        mb = B if S == 1 else max(1, B // max(1, num_minibatch or 1))      # M:1178
        L = self.config.num_hidden_layers
        if self.streamer is not None:
            self.streamer.begin()
        if spill is not None:
            spill.begin(pos0)
        for li in range(L):
            streamed = li >= self.n_resident
            v = self.streamer.acquire(li - self.n_resident) if streamed else self.resident_views[li]
            spilled = kcs[li] is None
            kc, vc = spill.acquire(li - (L - spill.n), pos0) if spilled else (kcs[li], vcs[li])
            for b0 in range(0, B, mb):
                nb = min(mb, B - b0)
                self.layer_rows(v, x[b0 * S:(b0 + nb) * S], kc, vc, nb, S, pos0, b0, ws)
            if spilled:
                spill.release(li - (L - spill.n), pos0, S)
            if streamed:
                self.streamer.release(li - self.n_resident)

    # ---- the reference's decoder face
    def forward(self, input_ids=None, attention_mask=None, head_mask=None, past_key_values=None, inputs_embeds=None,
                use_cache=None, output_attentions=None, output_hidden_states=None, return_dict=False,
                prefill_policy=None, decoding_policy=None, no_overlap=None, pin_weight=None, gpu_percentage=None,
                num_minibatch=None, enable_cxl=None, max_new_tokens=None):
        if head_mask is not None or inputs_embeds is not None or output_attentions or output_hidden_states:
            raise NotImplementedError("head_mask / inputs_embeds / output_attentions / output_hidden_states")
        _check_policy(prefill_policy, "prefill_policy")
        _check_policy(decoding_policy, "decoding_policy")
        self.check_gpu_percentage(gpu_percentage)
        ids = input_ids.to(self.device, torch.int64).contiguous()
        B, S = ids.shape
        past_len = int(past_key_values[0][0].shape[2]) if past_key_values is not None else 0   # M:1111
        if attention_mask is not None and attention_mask.shape[1] != past_len + S:
            raise ValueError(f"The provided attention mask has length {attention_mask.shape[1]}, but its length should "
                             f"be {past_len + S} (sum of the lengths of current and past inputs)")             # M:1127-1131
        cfg = self.config
        if S != 1:
            if past_len != 0:
                raise NotImplementedError("multi-token forward with a non-empty KV cache is not supported")
            new = max_new_tokens if max_new_tokens is not None else cfg.max_position_embeddings - S
            Hl = cfg.num_attention_heads // self.tp_world
            kcs = [torch.zeros(S + new, B, Hl, self.layout.dp, dtype=BF16, device=self.device) for _ in self.layers]
            vcs = [torch.zeros_like(k) for k in kcs]
            beam = torch.zeros(S + new, B, dtype=torch.long, device=self.device)
        else:
            kcs = [p[1] for p in past_key_values]
            vcs = [p[2] for p in past_key_values]
            beam = past_key_values[0][3]
        mb = B if S == 1 else max(1, B // max(1, num_minibatch or 1))
        ws = self.workspace(max(mb * S, B), B)
        am = None
        if attention_mask is not None:                 # only the learned positions depend on it (M:368-378; A:446-449, A:500)
            am = attention_mask.to(self.device, torch.int64).contiguous()
        x = torch.empty(B * S, cfg.hidden_size, dtype=BF16, device=self.device)
        self.set_overlap(not no_overlap)                                                                       # M:1173
        self.embed_rows(ids, past_len, am, x.view(B, S, cfg.hidden_size), ws)                                  # M:1107-1142
        self.run_layers(x, kcs, vcs, B, S, past_len, num_minibatch, ws)
        hidden = self.final_rows(x, ws).view(B, S, cfg.embed_dim)                                              # M:1563-1567
        T = past_len + S
        marker = torch.empty(1, T, T, 1, dtype=torch.long, device="meta")
        next_cache = tuple((marker, k, v, beam) for k, v in zip(kcs, vcs))
        return (hidden, next_cache)

    __call__ = forward


class _Model:
    def __init__(self, decoder):
        self.decoder = decoder


class OPTForCausalLM:
    """models.py:371-445 + the greedy loop of generation/greedy_search.py:37-460."""

    def __init__(self, config, device="cuda", tp_rank=0, tp_world=1):
        _lib.load()                                   # fail loudly if the CUDA library is missing
        self.config = config
        self.device = torch.device(device)
        self.tp_rank, self.tp_world = tp_rank, tp_world
        self.model = _Model(OPTDecoder(config, self.device, tp_rank, tp_world))
        self.layout = self.model.decoder.layout
        self._states = {}
        self.use_cuda_graphs = True
        self.kv_resident_layers = None      # None: as many layers' K/V in HBM as fit (rest spilled to pinned host)
        self.last_timing = None

    # ---- weights
    def init_weights(self, seed=0, kind="normal", gpu_percentage=100, bias_std=0.0, ln_std=0.0):
        """Random-init (lia/modeling_opt.py:895-904) or dummy (utils/opt-weight-gen.py:61-62) weights,
        generated layer by layer on the GPU from per-layer seeds (identical on every TP rank)."""
        cfg = self.config
        dec = self.model.decoder
        dec.load_embeddings(random_embeddings(cfg.vocab_size, cfg.hidden_size, cfg.max_position_embeddings,
                                              seed * 100003 + 17, self.device, kind, cfg.init_std, ln_std, cfg.pad_token_id,
                                              embed_dim=cfg.embed_dim, final_ln=cfg.do_layer_norm_before))
        dec.load_layers(lambda i, dev: random_layer(cfg.hidden_size, cfg.ffn_dim, seed * 100003 + 1000 + i, dev, kind,
                                                    cfg.init_std, bias_std, ln_std), gpu_percentage)
        return self

    def load_state_dict(self, sd, gpu_percentage=100):
        """HF OPT names (``model.decoder.layers.{i}.self_attn.q_proj.weight`` ...)."""
        dec = self.model.decoder
        dec.load_embeddings({"embed_tokens": sd["model.decoder.embed_tokens.weight"],
                             "embed_positions": sd["model.decoder.embed_positions.weight"],
                             "final_ln_w": sd.get("model.decoder.final_layer_norm.weight"),
                             "final_ln_b": sd.get("model.decoder.final_layer_norm.bias"),
                             "project_in": sd.get("model.decoder.project_in.weight"),
                             "project_out": sd.get("model.decoder.project_out.weight")})
        dec.load_layers(lambda i, dev: {k: t.to(dev, BF16) for k, t in layer_from_hf_state_dict(sd, i).items()},
                        gpu_percentage)
        return self

    @classmethod
    def from_pretrained(cls, path, device="cuda", gpu_percentage=100, tp_rank=0, tp_world=1, config=None, **unused):
        """Load a checkpoint directory (run_generation.py:159-167 ``from_pretrained``): either an HF OPT
        checkpoint (safetensors or pytorch_model.bin, sharded or not -- the form utils/opt-weight-gen.py:66-69
        writes) or this build's native slab directory (checkpoint.py), whose layers are read straight into
        HBM / the pinned arena without repacking."""
        from . import checkpoint
        ck = checkpoint.open_checkpoint(path)
        cfg = config or ck.config
        native = isinstance(ck, checkpoint.SlabCheckpoint)
        if native and ck.tp_world != tp_world:
            raise ValueError(f"{path} holds slabs for tensor-parallel world {ck.tp_world}, not {tp_world}; "
                             "re-run checkpoint.convert(..., tp_world=)")
        m = cls(cfg, device, tp_rank=tp_rank, tp_world=tp_world)
        dec = m.model.decoder
        dec.load_embeddings(ck.embeddings())
        if native:
            dec.load_packed(lambda i, out: ck.read_slab(i, tp_rank, out), gpu_percentage)
            ck.close()
        else:
            dec.load_layers(lambda i, dev: ck.layer(i, dev), gpu_percentage)
        return m

    def eval(self):
        return self

    # ---- reference forward face
    def forward(self, input_ids=None, attention_mask=None, past_key_values=None, head_mask=None, inputs_embeds=None,
                labels=None, use_cache=None, output_attentions=None, output_hidden_states=None, return_dict=False,
                prefill_policy=None, decoding_policy=None, no_overlap=None, pin_weight=None, gpu_percentage=None,
                num_minibatch=None, enable_cxl=None, max_new_tokens=None):
        if labels is not None:
            raise NotImplementedError("labels / loss: inference-only build")
        hidden, cache = self.model.decoder(
            input_ids=input_ids, attention_mask=attention_mask, head_mask=head_mask, past_key_values=past_key_values,
            inputs_embeds=inputs_embeds, use_cache=use_cache, output_attentions=output_attentions,
            output_hidden_states=output_hidden_states, prefill_policy=prefill_policy, decoding_policy=decoding_policy,
            no_overlap=no_overlap, pin_weight=pin_weight, gpu_percentage=gpu_percentage, num_minibatch=num_minibatch,
            enable_cxl=enable_cxl, max_new_tokens=max_new_tokens)
        if self.config.lm_head_generation and hidden.size(1) != 1:                       # models.py:424-431
            hidden = hidden[:, -1:, :]
        B, S, e = hidden.shape                                                            # e = word_embed_proj_dim
        logits = self.lm_head(hidden.reshape(B * S, e).contiguous()).view(B, S, self.config.vocab_size)
        return (logits, cache)

    __call__ = forward

    def lm_head(self, rows, out=None, ws=None):
        """Tied lm_head (lia/modeling_opt.py:1660), no bias: one GEMM against embed_tokens."""
        dec = self.model.decoder
        ws = ws or dec.workspace(rows.shape[0], rows.shape[0])
        return dec.k.gemm(rows, dec.embed_tokens, None, out=out, epilogue=EPI_BIAS, workspace=ws.gemm)

    # ---- generation
    def _state(self, B, S, new, num_minibatch):
        key = (B, S, new, num_minibatch, self.kv_resident_layers)
        if key not in self._states:
            for old in self._states.values():          # one live shape at a time: caches are large
                old.close()
            self._states.clear()
            self._states[key] = _GenState(self, B, S, new, num_minibatch)
        return self._states[key]

    def _head_and_pick(self, st, rows, t, suppress):
        dec = self.model.decoder
        xn = dec.final_rows(rows, st.ws, out=st.xn)                                       # M:1563-1567 (last position)
        self.lm_head(xn, out=st.logits, ws=st.ws)                                         # models.py:431
        self.model.decoder.k.argmax(st.logits, suppress, out=st.steps_tok[t])                              # greedy_search.py:367-395

    def _prefill(self, st, num_minibatch, suppress):
        dec, cfg = self.model.decoder, self.config
        B, S = st.B, st.S
        dec.embed_rows(st.prompt, 0, st.mask, st.x.view(B, S, cfg.hidden_size), st.ws)
        dec.run_layers(st.x, st.kc, st.vc, B, S, 0, num_minibatch, st.ws, st.spill)
        st.xd.copy_(st.x.view(B, S, cfg.hidden_size)[:, -1, :])                           # models.py:430 last token only
        self._head_and_pick(st, st.xd, 0, suppress)

    def _decode_step(self, st, t, suppress):
        """Generate token t (>= 1) from token t-1; cache holds S + t - 1 positions."""
        dec, cfg = self.model.decoder, self.config
        past = st.S + t - 1
        dec.embed_rows(st.steps_tok[t - 1].view(st.B, 1), past, st.mask, st.xd.view(st.B, 1, cfg.hidden_size), st.ws)
        dec.run_layers(st.xd, st.kc, st.vc, st.B, 1, past, 1, st.ws, st.spill)
        self._head_and_pick(st, st.xd, t, suppress)

    def _decode_program(self, st):
        """The decode step of this generation shape as ONE persistent kernel (program.py, csrc/decode_program_sm100.cu),
        built by running ``_decode_step`` once against a recorder instead of the kernel layer.  None when the step cannot be
        a program (not asked for with LIA_DECODE_PROGRAM=1 -- the CUDA-graph replay of the kernel-per-operation step is the
        faster of the two today, profiles/README.md --, more than 128 sequences, a host-side collective in the step, too many cached
        positions for the score buffer): the caller then replays CUDA graphs of the kernel-per-operation step."""
        if st.program is None:
            st.program = False
            if os.environ.get("LIA_DECODE_PROGRAM", "0") != "0" and st.B <= 128 and st.new > 1 and self.device.type == "cuda":
                dec = self.model.decoder
                prog = None
                try:
                    prog = program_mod.DecodeProgram(st.B)
                    dec.k = program_mod.ProgramRecorder(prog, self.device)
                    self._decode_step(st, 1, -1)
                    prog.finalize()
                    st.program = prog
                except program_mod.ProgramUnsupported:
                    if prog is not None:
                        prog.close()
                finally:
                    dec.k = ops
        return st.program or None

    def generate(self, input_ids, max_new_tokens=32, min_new_tokens=None, do_sample=False, num_beams=1,
                 temperature=None, attention_mask=None, prefill_policy=None, decoding_policy=None, no_overlap=None,
                 pin_weight=None, gpu_percentage=None, num_minibatch=None, enable_cxl=None, **unused):
        """Greedy generation (run_generation.py:179-182: do_sample=False, num_beams=1,
        min_new_tokens == max_new_tokens).  Returns ids [B, S+new] on ``input_ids.device`` -- and the
        per-token latency list when ``config.token_latency`` (greedy_search.py:455-456).  Length is
        the only stop criterion (greedy_search.py:425); eos is suppressed while fewer than
        ``min_new_tokens`` tokens exist (generation_utils.py:872-880)."""
        if do_sample or num_beams != 1:
            raise NotImplementedError("only greedy search (do_sample=False, num_beams=1) is implemented")
        _check_policy(prefill_policy, "prefill_policy")
        _check_policy(decoding_policy, "decoding_policy")
        dec = self.model.decoder
        dec.check_gpu_percentage(gpu_percentage)
        num_minibatch = int(num_minibatch or 1)
        B, S = input_ids.shape
        new = int(max_new_tokens)
        if new < 1:
            raise ValueError(f"max_new_tokens must be at least 1 (got {max_new_tokens})")
        if S + new > self.config.max_position_embeddings:
            raise ValueError(f"S + max_new_tokens = {S + new} exceeds max_position_embeddings")
        min_new = int(min_new_tokens or 0)
        st = self._state(B, S, new, num_minibatch)
        st.calls += 1
        dec.set_overlap(not no_overlap, st.spill)                                        # M:1173
        eos = self.config.eos_token_id
        graphs_ok = self.use_cuda_graphs and dec.streamer is None and st.spill is None
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(new + 1)]
        st.prompt.copy_(input_ids, non_blocking=True)                                    # H2D (pinned host -> HBM)
        pad = self.config.pad_token_id
        if attention_mask is not None:
            if tuple(attention_mask.shape) != (B, S):
                raise ValueError(f"attention_mask must be [{B}, {S}], got {tuple(attention_mask.shape)}")
            st.mask[:, :S].copy_(attention_mask, non_blocking=True)
        elif pad is not None and pad != eos:
            # generation_utils.py:469-485: without a mask, pad ids in the prompt define one (all ones when there are none)
            st.mask[:, :S].copy_(st.prompt.ne(pad))
        else:
            st.mask[:, :S].fill_(1)
        ev[0].record()
        self._prefill(st, num_minibatch, eos if 0 < min_new else -1)
        ev[1].record()
        prog = self._decode_program(st) if graphs_ok else None
        for t in range(1, new):
            suppress = eos if t < min_new else -1
            if prog is not None:
                # one persistent kernel per step: embedding, every layer, final LayerNorm, lm_head and argmax
                prog.run(st.S + t - 1, st.steps_tok[t - 1], st.steps_tok[t], suppress)
            elif graphs_ok and st.calls >= 2:
                g = st.graphs.get((t, suppress))
                if g is None:
                    g = torch.cuda.CUDAGraph()
                    with graphs.capture(g):              # GC-quiesced: a finalizer must not invalidate the capture
                        self._decode_step(st, t, suppress)
                    st.graphs[(t, suppress)] = g
                g.replay()
            else:
                self._decode_step(st, t, suppress)
            ev[t + 1].record()
        out_dev = torch.cat([st.prompt, st.steps_tok[:new].t()], dim=1)
        out = out_dev.to(input_ids.device)                                               # D2H of the result (syncs)
        torch.cuda.synchronize(self.device)
        if st.arena is not None and os.environ.get("LIA_TP_SELF_LOOP", "0") == "0":
            st.arena.check()                          # a peer that never showed up: raise instead of returning garbage
            if st.arena_d is not st.arena:
                st.arena_d.check()
        if prog is not None:
            prog.check()                              # a CTA that timed out on a dependency: raise, do not return garbage
        lat = [ev[i].elapsed_time(ev[i + 1]) / 1e3 for i in range(new)]
        self.last_timing = {"prefill_s": lat[0], "decode_s": lat[1:], "total_s": sum(lat)}
        if self.config.token_latency:
            return out, lat
        return out
