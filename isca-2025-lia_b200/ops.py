"""Tensor-level wrappers over the C ABI (include/lia_b200.h).

PyTorch is plumbing here: it owns device memory and streams; every op below is one
call into libliab200.so on ``torch.cuda.current_stream()``.  No op has a PyTorch
fallback -- a missing library or a non-CUDA tensor raises.
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import EPI_BIAS, EPI_BIAS_RELU, EPI_BIAS_RESIDUAL, EPI_QKV, LiaQkvArgs, check  # noqa: F401

BF16 = torch.bfloat16


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return t.data_ptr() if t is not None else None


def _req(t, name, dtype=BF16):
    if not (t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise _lib.LiaError(f"{name}: need a contiguous CUDA {dtype} tensor, got {t.dtype} on {t.device} "
                            f"contiguous={t.is_contiguous()}")
    return t


def count_launches(n=1):
    _lib.launch_count += n


def layernorm(x, w, b, eps=1e-5, out=None):
    """F.layer_norm over the last dim (decoder.py:107-119)."""
    _req(x, "x"); _req(w, "w"); _req(b, "b")
    h = x.shape[-1]
    rows = x.numel() // h
    if out is None:
        out = torch.empty_like(x)
    check(_lib.load().lia_layernorm_bf16(_p(x), _p(w), _p(b), _p(_req(out, "out")), rows, h, eps, _stream()),
          "lia_layernorm_bf16")
    count_launches()
    return out


class GemmWorkspace:
    """Split-K workspace, zero-initialised once (the kernel leaves its counters zeroed)."""

    def __init__(self, nbytes, device):
        self.buf = torch.zeros(max(int(nbytes), 16384), dtype=torch.uint8, device=device)

    @staticmethod
    def bytes_for(shapes):
        lib = _lib.load()
        return max([lib.lia_gemm_workspace_bytes(m, n, k) for (m, n, k) in shapes] + [16384])


def gemm(a, w, bias, out=None, epilogue=EPI_BIAS, residual=None, qkv=None, workspace=None):
    """out = epilogue(a @ w.T) -- a [M,K], w [N,K] (decoder.py:79-105, attentions.py:376-418)."""
    _req(a, "a"); _req(w, "w")
    M, K = a.shape
    N = w.shape[0]
    if w.shape[1] != K:
        raise _lib.LiaError(f"gemm: a is [{M},{K}] but w is {tuple(w.shape)}")
    if bias is not None:
        _req(bias, "bias")
    if residual is not None:
        _req(residual, "residual")
    if epilogue != EPI_QKV and out is None:
        out = torch.empty(M, N, dtype=BF16, device=a.device)
    if epilogue != EPI_QKV:
        _req(out, "out")
        if tuple(out.shape) != (M, N):
            raise _lib.LiaError(f"gemm: out is {tuple(out.shape)}, need ({M}, {N})")
    if residual is not None and tuple(residual.shape) != (M, N):
        raise _lib.LiaError(f"gemm: residual is {tuple(residual.shape)}, need ({M}, {N})")
    ws_ptr, ws_bytes = (None, 0)
    if workspace is not None:
        ws_ptr, ws_bytes = workspace.buf.data_ptr(), workspace.buf.numel()
    check(_lib.load().lia_gemm_bf16(_p(a), _p(w), _p(bias), _p(residual), _p(out), M, N, K, epilogue,
                                    ctypes.byref(qkv) if qkv is not None else None, ws_ptr, ws_bytes, _stream()),
          "lia_gemm_bf16")
    count_launches()
    return out


def gemm_allreduce(a, w, bias, residual, out, tp_args, workspace=None):
    """Row-parallel projection fused with its all-reduce and residual add (world > 1):
    out = residual + sum_over_ranks(a_r @ w_r.T + bias_r)  (decoder.py:60-77 + 247/317).
    ``tp_args`` is a LiaTpArgs from tp.PeerArena.args(); for M > 128 ``out`` must be an arena tensor."""
    _req(a, "a"); _req(w, "w"); _req(residual, "residual"); _req(out, "out")
    M, K = a.shape
    N = w.shape[0]
    if w.shape[1] != K:
        raise _lib.LiaError(f"gemm_allreduce: a is [{M},{K}] but w is {tuple(w.shape)}")
    if tuple(out.shape) != (M, N) or tuple(residual.shape) != (M, N):
        raise _lib.LiaError(f"gemm_allreduce: out {tuple(out.shape)} / residual {tuple(residual.shape)}, need ({M}, {N})")
    if bias is not None:
        _req(bias, "bias")
    ws_ptr, ws_bytes = (None, 0)
    if workspace is not None:
        ws_ptr, ws_bytes = workspace.buf.data_ptr(), workspace.buf.numel()
    check(_lib.load().lia_gemm_allreduce_bf16(_p(a), _p(w), _p(bias), _p(residual), _p(out), M, N, K,
                                              ctypes.byref(tp_args), ws_ptr, ws_bytes, _stream()),
          "lia_gemm_allreduce_bf16")
    count_launches()
    return out


def _check_cache_rows(k_cache, v_cache, rows, what):
    """The time dimension of the cache bounds every append and every read (the reference's slice assignment
    A:473-491 raises there)."""
    if k_cache.dim() != 4 or k_cache.shape != v_cache.shape:
        raise _lib.LiaError(f"{what}: k_cache {tuple(k_cache.shape)} / v_cache {tuple(v_cache.shape)} must be equal [Tmax,B,H,d]")
    if rows > k_cache.shape[0]:
        raise _lib.LiaError(f"{what}: {rows} cached positions exceed the cache's {k_cache.shape[0]} rows")


def qkv_args(q_out, k_cache, v_cache, S, pos0, b0, scale):
    """k_cache/v_cache: [Tmax, Bc, H, d] time-major (attentions.py:471-472)."""
    _req(q_out, "q_out"); _req(k_cache, "k_cache"); _req(v_cache, "v_cache")
    hq = k_cache.shape[2] * k_cache.shape[3]
    _check_cache_rows(k_cache, v_cache, pos0 + S, "qkv_args")
    return LiaQkvArgs(q_out.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(), hq, S, pos0, k_cache.shape[1], b0,
                      float(scale))


def kv_append(q, k, v, k_cache, v_cache, pos0, b0, scale, q_out=None):
    """q_out = q*scale; append k/v rows [B,S,H,d] to the time-major cache (attentions.py:456-491)."""
    _req(q, "q"); _req(k, "k"); _req(v, "v"); _req(k_cache, "k_cache"); _req(v_cache, "v_cache")
    B, S = q.shape[0], q.shape[1]
    hq = k_cache.shape[2] * k_cache.shape[3]
    _check_cache_rows(k_cache, v_cache, pos0 + S, "kv_append")
    if q_out is None:
        q_out = torch.empty_like(q)
    check(_lib.load().lia_kv_append_bf16(_p(q), _p(k), _p(v), _p(q_out), _p(k_cache), _p(v_cache), B, S, hq, pos0,
                                         k_cache.shape[1], b0, float(scale), _stream()), "lia_kv_append_bf16")
    count_launches()
    return q_out


def attn_prefill(q, k_cache, v_cache, B, S, b0=0, out=None):
    """Causal attention over rows [0,S) of the cache (attentions.py:444-449, 493-536)."""
    _req(q, "q"); _req(k_cache, "k_cache"); _req(v_cache, "v_cache")
    _, Bc, H, d = k_cache.shape
    _check_cache_rows(k_cache, v_cache, S, "attn_prefill")
    if out is None:
        out = torch.empty(B * S, H * d, dtype=BF16, device=q.device)
    check(_lib.load().lia_attn_prefill_bf16(_p(q), _p(k_cache), _p(v_cache), _p(_req(out, "out")), B, H, S, d, Bc, b0,
                                            _stream()), "lia_attn_prefill_bf16")
    count_launches()
    return out


_SM_COUNT = {}


def _sm_count(device):
    key = torch.device(device).index or 0
    if key not in _SM_COUNT:
        _SM_COUNT[key] = torch.cuda.get_device_properties(key).multi_processor_count
    return _SM_COUNT[key]


def attn_decode_launches(B, H, T, splits, has_workspace, sms):
    """Kernels one lia_attn_decode_bf16 call launches: 1, or 2 when it splits the keys (the split kernel and the combine).
    Bookkeeping for ``gpu_launches`` only -- restates the choice made in csrc/attn_decode.cu (pick_splits and the fallbacks
    after it); the kernel never sees this."""
    fewest = (T + 4095) // 4096
    if splits <= 0:
        per_sm = -splits
        if per_sm == 0:
            env = os.environ.get("LIA_ATTN_CTAS_PER_SM", "")
            per_sm = int(env) if env.isdigit() and int(env) > 0 else 2
        want = (per_sm * sms + B * H - 1) // (B * H)
        splits = max(min(want, max(T // 128, 1), 32), fewest)
    if splits > 1 and not has_workspace:
        splits = fewest
    t_chunk = (T + splits - 1) // splits
    return 2 if (T + t_chunk - 1) // t_chunk > 1 else 1


def attn_decode(q, k_cache, v_cache, B, T, b0=0, out=None, splits=0, workspace=None):
    """One query token per sequence over T cached positions, no mask (attentions.py:395-399, 500)."""
    _req(q, "q"); _req(k_cache, "k_cache"); _req(v_cache, "v_cache")
    _, Bc, H, d = k_cache.shape
    _check_cache_rows(k_cache, v_cache, T, "attn_decode")
    if out is None:
        out = torch.empty(B, H * d, dtype=BF16, device=q.device)
    ws_ptr, ws_bytes = (None, 0)
    if workspace is not None:
        ws_ptr, ws_bytes = workspace.data_ptr(), workspace.numel() * workspace.element_size()
    check(_lib.load().lia_attn_decode_bf16(_p(q), _p(k_cache), _p(v_cache), _p(_req(out, "out")), B, H, T, d, Bc, b0,
                                           splits, ws_ptr, ws_bytes, _stream()), "lia_attn_decode_bf16")
    count_launches(attn_decode_launches(B, H, T, splits, workspace is not None, _sm_count(q.device)))
    return out


def attn_decode_workspace(B, H, d, device, max_splits=32):
    n = _lib.load().lia_attn_decode_workspace_bytes(B, H, d, max_splits)
    return torch.empty(n // 4, dtype=torch.float32, device=device)


def embed(ids, embed_tokens, embed_positions, past_len, out=None, attention_mask=None):
    """Token + learned positional embedding (lia/modeling_opt.py:1107-1142, 368-378).  ``attention_mask``
    (int64 [B, >= past_len + S], may be a column-slice view of a wider buffer) selects the positions the
    reference derives from its cumsum; None means all ones."""
    _req(ids, "ids", torch.int64)
    if embed_tokens is None and embed_positions is None:
        raise _lib.LiaError("embed: at most one of embed_tokens / embed_positions may be None")
    if embed_tokens is not None:
        _req(embed_tokens, "embed_tokens")
    if embed_positions is not None:
        _req(embed_positions, "embed_positions")
    B, S = ids.shape
    V = embed_tokens.shape[0] if embed_tokens is not None else 0
    h = (embed_tokens if embed_tokens is not None else embed_positions).shape[1]
    if embed_tokens is not None and embed_positions is not None and embed_positions.shape[1] != h:
        raise _lib.LiaError(f"embed: token rows are {h} wide but position rows {embed_positions.shape[1]}")
    P = embed_positions.shape[0] if embed_positions is not None else 0
    if out is not None and (out.numel() != B * S * h):
        raise _lib.LiaError(f"embed: out has {out.numel()} elements, need {B * S * h}")
    if out is None:
        out = torch.empty(B, S, h, dtype=BF16, device=ids.device)
    mask_ptr, mask_ld = None, 0
    if attention_mask is not None:
        am = attention_mask
        if not (am.is_cuda and am.dtype == torch.int64 and am.dim() == 2 and am.shape[0] == B and am.stride(1) == 1):
            raise _lib.LiaError(f"attention_mask: need a CUDA int64 [B, T] tensor with unit column stride, got {am.dtype} "
                                f"{tuple(am.shape)} strides {am.stride()} on {am.device}")
        if am.shape[1] < past_len + S:
            raise _lib.LiaError(f"attention_mask has {am.shape[1]} columns, need past_len + S = {past_len + S}")
        mask_ptr, mask_ld = am.data_ptr(), am.stride(0) if B > 1 else max(am.stride(0), am.shape[1])
    check(_lib.load().lia_embed_masked_bf16(_p(ids), mask_ptr, mask_ld, _p(embed_tokens), _p(embed_positions),
                                            _p(_req(out, "out")), B, S, h, past_len, V, P, _stream()),
          "lia_embed_masked_bf16")
    count_launches()
    return out


def argmax(logits, suppress_id=-1, out=None):
    """Greedy pick with one suppressed id (generation_utils.py:872-880, greedy_search.py:395)."""
    _req(logits, "logits")
    B, V = logits.shape
    if out is None:
        out = torch.empty(B, dtype=torch.int64, device=logits.device)
    check(_lib.load().lia_argmax_bf16(_p(logits), _p(_req(out, "out", torch.int64)), B, V, suppress_id, _stream()),
          "lia_argmax_bf16")
    count_launches()
    return out


def residual_add(x, residual, out=None):
    """out = residual + x after a tensor-parallel all-reduce (decoder.py:247, 317)."""
    _req(x, "x"); _req(residual, "residual")
    if out is None:
        out = torch.empty_like(x)
    check(_lib.load().lia_residual_add_bf16(_p(x), _p(residual), _p(_req(out, "out")), x.numel(), _stream()),
          "lia_residual_add_bf16")
    count_launches()
    return out
