"""On-disk weights -> per-layer slabs (the weight file format either side of the path).

Two on-disk forms are read:

1. **HF checkpoint directory** -- what ``from_pretrained`` consumes at
   single_instance/run_generation.py:159-167 and what the reference's dummy-weight generator writes
   (examples/cpu/inference/python/llm/utils/opt-weight-gen.py:66-69, ``save_pretrained(...,
   safe_serialization=False)``): ``config.json`` plus one of ``model.safetensors``,
   ``model.safetensors.index.json`` + shards, ``pytorch_model.bin``, ``pytorch_model.bin.index.json`` +
   shards.  Tensors are fetched lazily, one layer at a time (a 175B checkpoint is never resident in
   host memory as a whole), converted to bf16 and packed into the layer slab (weights.pack_layer).
   The safetensors container is parsed here (8-byte little-endian header length, JSON header, raw
   little-endian tensor bytes) -- no dependency on the ``safetensors`` package.

2. **Native slab directory** (``lia_slabs.json`` + ``rank{r}.slabs`` + ``embeddings.bin``): every layer is
   stored exactly as it sits in HBM or in the pinned host arena -- this rank's TP shard, q/k/v fused
   row-wise, bf16, ``LayerLayout`` offsets -- so loading a streamed layer is ONE ``readinto`` straight
   into the arena the streamer copies from, with no repacking and no intermediate tensor.  This is the
   format the layer streamer consumes; ``convert()`` writes it from an HF checkpoint and
   ``write_slabs()`` from any layer generator (``scripts/opt_weight_gen.py``: the reference's dummy
   66B/175B weights without ever materialising the model).
"""
import json
import os
import struct

import numpy as np
import torch

from .weights import LAYER_KEYS, LayerLayout, layer_from_hf_state_dict, pack_layer

BF16 = torch.bfloat16
SLAB_FORMAT = "lia-b200-slabs"
SLAB_VERSION = 1
_ST_DTYPES = {"BF16": (torch.bfloat16, 2), "F16": (torch.float16, 2), "F32": (torch.float32, 4), "F64": (torch.float64, 8),
              "I64": (torch.int64, 8), "I32": (torch.int32, 4), "I16": (torch.int16, 2), "I8": (torch.int8, 1),
              "U8": (torch.uint8, 1), "BOOL": (torch.bool, 1)}
_CONFIG_FIELDS = ("hidden_size", "num_hidden_layers", "num_attention_heads", "ffn_dim", "vocab_size",
                  "max_position_embeddings", "do_layer_norm_before", "word_embed_proj_dim", "pad_token_id", "bos_token_id",
                  "eos_token_id", "init_std")
_EMBED_KEYS = {"embed_tokens": "model.decoder.embed_tokens.weight", "embed_positions": "model.decoder.embed_positions.weight",
               "final_ln_w": "model.decoder.final_layer_norm.weight", "final_ln_b": "model.decoder.final_layer_norm.bias",
               "project_in": "model.decoder.project_in.weight", "project_out": "model.decoder.project_out.weight"}


class CheckpointError(RuntimeError):
    pass


# ------------------------------------------------------------------------------------------------
# safetensors container
# ------------------------------------------------------------------------------------------------
class SafetensorsFile:
    """Lazy reader of one ``.safetensors`` file: header parsed at open, tensor bytes memory-mapped."""

    def __init__(self, path):
        self.path = path
        size = os.path.getsize(path)
        with open(path, "rb") as f:
            head = f.read(8)
            if len(head) != 8:
                raise CheckpointError(f"{path}: truncated safetensors header")
            (n,) = struct.unpack("<Q", head)
            if n <= 0 or 8 + n > size:
                raise CheckpointError(f"{path}: header length {n} exceeds the file size {size}")
            try:
                self.header = json.loads(f.read(n).decode("utf-8"))
            except (UnicodeDecodeError, json.JSONDecodeError) as e:
                raise CheckpointError(f"{path}: malformed safetensors header ({e})") from None
        self.metadata = self.header.pop("__metadata__", {})
        self.data_start = 8 + n
        self.data_bytes = size - self.data_start
        for k, e in self.header.items():
            b, en = e["data_offsets"]
            if e["dtype"] not in _ST_DTYPES:
                raise CheckpointError(f"{path}: tensor {k!r} has unsupported dtype {e['dtype']}")
            want = int(np.prod(e["shape"], dtype=np.int64)) * _ST_DTYPES[e["dtype"]][1]
            if not (0 <= b <= en <= self.data_bytes) or en - b != want:
                raise CheckpointError(f"{path}: tensor {k!r} has inconsistent offsets {b}:{en} for shape {e['shape']}")
        self._mm = None

    def keys(self):
        return self.header.keys()

    def get(self, key):
        """CPU tensor in the file's dtype (a private copy: the mapping may be closed afterwards)."""
        e = self.header[key]
        dt, _ = _ST_DTYPES[e["dtype"]]
        b, en = e["data_offsets"]
        if self._mm is None:
            self._mm = np.memmap(self.path, dtype=np.uint8, mode="r", offset=self.data_start) if self.data_bytes else \
                np.zeros(0, dtype=np.uint8)
        raw = torch.from_numpy(np.array(self._mm[b:en]))        # copy out of the mapping (aligned, owned)
        if raw.numel() == 0:
            return torch.empty(e["shape"], dtype=dt)
        return raw.view(dt).reshape(e["shape"])

    def close(self):
        self._mm = None


def write_safetensors(path, tensors, metadata=None):
    """Minimal writer (tests, tooling): ``tensors`` name -> CPU tensor."""
    rev = {v[0]: k for k, v in _ST_DTYPES.items()}
    header, off, blobs = {}, 0, []
    for k in sorted(tensors):
        t = tensors[k].detach().contiguous().cpu()
        raw = t.reshape(-1).view(torch.uint8) if t.numel() else torch.zeros(0, dtype=torch.uint8)
        header[k] = {"dtype": rev[t.dtype], "shape": list(t.shape), "data_offsets": [off, off + raw.numel()]}
        off += raw.numel()
        blobs.append(raw.numpy().tobytes())
    if metadata:
        header["__metadata__"] = {str(k): str(v) for k, v in metadata.items()}
    hb = json.dumps(header, separators=(",", ":")).encode("utf-8")
    hb += b" " * (-len(hb) % 8)                                  # keep the data section 8-byte aligned
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hb)))
        f.write(hb)
        for b in blobs:
            f.write(b)


# ------------------------------------------------------------------------------------------------
# HF checkpoint directory
# ------------------------------------------------------------------------------------------------
def _norm_key(k):
    """facebook/opt-* hub checkpoints omit the leading ``model.``; normalise to the names of SURVEY.md 8a0."""
    if k.startswith("decoder."):
        return "model." + k
    return k


class _LazyStateDict:
    """Mapping name -> tensor over the shard files of an HF checkpoint; at most one ``.bin`` shard
    (torch.load with mmap) or the safetensors headers are held at a time."""

    def __init__(self, directory):
        self.dir = directory
        self.where = {}        # normalised key -> (file, key in file)
        self.kind = None
        j = os.path.join
        if os.path.exists(j(directory, "model.safetensors.index.json")):
            self.kind = "safetensors"
            wm = json.load(open(j(directory, "model.safetensors.index.json")))["weight_map"]
            for k, fn in wm.items():
                self.where[_norm_key(k)] = (fn, k)
        elif os.path.exists(j(directory, "model.safetensors")):
            self.kind = "safetensors"
            for k in SafetensorsFile(j(directory, "model.safetensors")).keys():
                self.where[_norm_key(k)] = ("model.safetensors", k)
        elif os.path.exists(j(directory, "pytorch_model.bin.index.json")):
            self.kind = "bin"
            wm = json.load(open(j(directory, "pytorch_model.bin.index.json")))["weight_map"]
            for k, fn in wm.items():
                self.where[_norm_key(k)] = (fn, k)
        elif os.path.exists(j(directory, "pytorch_model.bin")):
            self.kind = "bin"
            sd = self._load_bin("pytorch_model.bin")
            for k in sd:
                self.where[_norm_key(k)] = ("pytorch_model.bin", k)
        else:
            raise CheckpointError(f"{directory}: no model.safetensors[.index.json] or pytorch_model.bin[.index.json]")
        self._st = {}

    _bin_name, _bin_sd = None, None

    def _load_bin(self, fn):
        if self._bin_name != fn:
            self._bin_sd = None
            path = os.path.join(self.dir, fn)
            try:
                self._bin_sd = torch.load(path, map_location="cpu", weights_only=True, mmap=True)
            except (RuntimeError, ValueError):                    # legacy (non-zip) pickles cannot be mmapped
                self._bin_sd = torch.load(path, map_location="cpu", weights_only=True)
            self._bin_name = fn
        return self._bin_sd

    def __contains__(self, k):
        return k in self.where

    def __getitem__(self, k):
        if k not in self.where:
            raise KeyError(f"{k} not found in checkpoint {self.dir}")
        fn, fk = self.where[k]
        if self.kind == "safetensors":
            if fn not in self._st:
                self._st[fn] = SafetensorsFile(os.path.join(self.dir, fn))
            return self._st[fn].get(fk)
        return self._load_bin(fn)[fk]

    def keys(self):
        return self.where.keys()


def config_from_json(path):
    """``config.json`` -> OPTConfig with the fields the path reads (lia/modeling_opt.py:977-1013)."""
    from .modeling_opt import OPTConfig
    c = json.load(open(path))
    if c.get("model_type", "opt") != "opt":
        raise CheckpointError(f"{path}: model_type {c.get('model_type')!r} is not OPT")
    if c.get("_remove_final_layer_norm", False):
        raise NotImplementedError("_remove_final_layer_norm (pre-v4.20.1 fine-tunes, lia/modeling_opt.py:998-1001)")
    if c.get("activation_function", "relu") != "relu":
        raise NotImplementedError(f"activation {c['activation_function']!r}: OPT uses ReLU (decoder.py:92-105)")
    kw = {k: c[k] for k in _CONFIG_FIELDS if k in c and c[k] is not None}
    if kw.get("word_embed_proj_dim") == c["hidden_size"]:
        kw["word_embed_proj_dim"] = 0                          # 0 = "same as hidden_size"
    name = c.get("_name_or_path") or os.path.basename(os.path.dirname(os.path.abspath(path)))
    return OPTConfig(name=str(name).rstrip("/").split("/")[-1] or "opt", **kw)


class HFCheckpoint:
    """HF checkpoint directory as a per-layer source."""

    def __init__(self, directory):
        self.dir = directory
        cj = os.path.join(directory, "config.json")
        if not os.path.exists(cj):
            raise CheckpointError(f"{directory}: config.json not found")
        self.config = config_from_json(cj)
        self.sd = _LazyStateDict(directory)

    def layer(self, i, device="cpu"):
        """Full (unsharded) layer dict in bf16 on ``device`` (keys weights.LAYER_KEYS)."""
        w = layer_from_hf_state_dict(self.sd, i)
        h, f = self.config.hidden_size, self.config.ffn_dim
        want = {"q_w": (h, h), "k_w": (h, h), "v_w": (h, h), "o_w": (h, h), "fc1_w": (f, h), "fc2_w": (h, f), "fc1_b": (f,)}
        for k, shp in want.items():
            if tuple(w[k].shape) != shp:
                raise CheckpointError(f"layer {i} tensor {k} has shape {tuple(w[k].shape)}, config says {shp}")
        return {k: t.to(device=device, dtype=BF16) for k, t in w.items()}

    def embeddings(self):
        sd, cfg = self.sd, self.config
        # final LayerNorm only for pre-LN models, project_in/out only when the token table is narrower (opt-350m)
        e = {k: sd[name] for k, name in _EMBED_KEYS.items() if name in sd}
        V, P = cfg.vocab_size, cfg.max_position_embeddings + 2                            # offset 2, M:365-366
        if e["embed_tokens"].shape[0] != V or e["embed_positions"].shape[0] != P:
            raise CheckpointError(f"embedding tables {tuple(e['embed_tokens'].shape)} / {tuple(e['embed_positions'].shape)} "
                                  f"do not match config (vocab {V}, positions {P})")
        return {k: t.to(BF16) for k, t in e.items()}


# ------------------------------------------------------------------------------------------------
# native slab directory
# ------------------------------------------------------------------------------------------------
def _flat_bytes(t):
    """uint8 numpy view of a contiguous CPU tensor (no copy)."""
    return t.reshape(-1).view(torch.uint8).numpy()


def write_slabs(directory, config, layer_fn, embeddings, tp_world=1):
    """Write the native format.  ``layer_fn(i)`` -> full layer dict (any float dtype, CPU or GPU);
    one layer is alive at a time.  Every rank's shard file is written in the same pass."""
    os.makedirs(directory, exist_ok=True)
    layout = LayerLayout(config.hidden_size, config.ffn_dim, tp_world, heads=config.num_attention_heads)
    files = [open(os.path.join(directory, f"rank{r}.slabs"), "wb") for r in range(tp_world)]
    try:
        for i in range(config.num_hidden_layers):
            w = {k: t.to(BF16) for k, t in layer_fn(i).items()}
            if set(w) != set(LAYER_KEYS):
                raise CheckpointError(f"layer {i}: expected keys {LAYER_KEYS}")
            for r in range(tp_world):
                slab = pack_layer(w, layout, r).cpu().contiguous()
                files[r].write(_flat_bytes(slab).tobytes())
    finally:
        for f in files:
            f.close()
    order = [k for k in _EMBED_KEYS if embeddings.get(k) is not None]
    emb = {}
    off = 0
    with open(os.path.join(directory, "embeddings.bin"), "wb") as f:
        for k in order:
            t = embeddings[k].to(BF16).cpu().contiguous()
            emb[k] = {"shape": list(t.shape), "offset": off}
            f.write(_flat_bytes(t).tobytes())
            off += t.numel() * 2
    meta = {"format": SLAB_FORMAT, "version": SLAB_VERSION, "dtype": "bf16", "tp_world": tp_world,
            "config": {k: getattr(config, k) for k in _CONFIG_FIELDS}, "name": config.name,
            "layout": {"numel": layout.numel, "offsets": {k: list(v) for k, v in layout.offsets.items()},
                       "shapes": {k: list(v) for k, v in layout.shapes.items()}},
            "embeddings": emb}
    with open(os.path.join(directory, "lia_slabs.json"), "w") as f:
        json.dump(meta, f, indent=1)
    return meta


class SlabCheckpoint:
    """Native slab directory: ``read_slab(i, rank, out)`` fills ``out`` (flat CPU bf16, e.g. a slice of
    the pinned host arena) with one ``readinto``."""

    def __init__(self, directory):
        self.dir = directory
        from .modeling_opt import OPTConfig
        m = json.load(open(os.path.join(directory, "lia_slabs.json")))
        if m.get("format") != SLAB_FORMAT or m.get("version") != SLAB_VERSION or m.get("dtype") != "bf16":
            raise CheckpointError(f"{directory}: not a {SLAB_FORMAT} v{SLAB_VERSION} bf16 directory")
        self.meta = m
        self.tp_world = int(m["tp_world"])
        self.config = OPTConfig(name=m.get("name", "opt"), **m["config"])
        self.layout = LayerLayout(self.config.hidden_size, self.config.ffn_dim, self.tp_world, heads=self.config.num_attention_heads)
        if self.layout.numel != m["layout"]["numel"] or {k: list(v) for k, v in self.layout.offsets.items()} != m["layout"]["offsets"]:
            raise CheckpointError(f"{directory}: slab layout in the file differs from this build's LayerLayout")
        for r in range(self.tp_world):
            p = os.path.join(directory, f"rank{r}.slabs")
            want = self.layout.nbytes * self.config.num_hidden_layers
            if not os.path.exists(p) or os.path.getsize(p) != want:
                raise CheckpointError(f"{p}: missing or not {want} bytes")
        self._f = {}

    def read_slab(self, i, rank, out):
        if not (0 <= i < self.config.num_hidden_layers and 0 <= rank < self.tp_world):
            raise IndexError((i, rank))
        if out.dtype != BF16 or out.device.type != "cpu" or not out.is_contiguous() or out.numel() != self.layout.numel:
            raise CheckpointError("read_slab: `out` must be a contiguous CPU bf16 tensor of layout.numel elements")
        f = self._f.get(rank)
        if f is None:
            f = self._f[rank] = open(os.path.join(self.dir, f"rank{rank}.slabs"), "rb", buffering=0)
        f.seek(i * self.layout.nbytes)
        buf = memoryview(_flat_bytes(out))
        got = 0
        while got < self.layout.nbytes:
            n = f.readinto(buf[got:])
            if not n:
                raise CheckpointError(f"short read of layer {i} from {self.dir}/rank{rank}.slabs")
            got += n
        return out

    def embeddings(self):
        out = {}
        with open(os.path.join(self.dir, "embeddings.bin"), "rb") as f:
            for k, e in self.meta["embeddings"].items():
                n = int(np.prod(e["shape"], dtype=np.int64))
                f.seek(e["offset"])
                raw = np.frombuffer(f.read(n * 2), dtype=np.uint8).copy()
                if raw.size != n * 2:
                    raise CheckpointError(f"{self.dir}/embeddings.bin: short read of {k}")
                out[k] = torch.from_numpy(raw).view(BF16).view(*e["shape"])
        return out

    def close(self):
        for f in self._f.values():
            f.close()
        self._f = {}


def open_checkpoint(path):
    """Directory -> SlabCheckpoint (native) or HFCheckpoint."""
    if not os.path.isdir(path):
        raise CheckpointError(f"{path}: not a directory")
    if os.path.exists(os.path.join(path, "lia_slabs.json")):
        return SlabCheckpoint(path)
    return HFCheckpoint(path)


def convert(src, dst, tp_world=1):
    """HF checkpoint directory -> native slab directory for ``tp_world`` ranks."""
    ck = HFCheckpoint(src)
    return write_slabs(dst, ck.config, lambda i: ck.layer(i), ck.embeddings(), tp_world)
