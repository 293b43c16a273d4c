"""Decode programs: a whole decode step as ONE persistent kernel (csrc/decode_program_sm100.cu).

``DecodeProgram`` wraps the C ABI (include/lia_b200.h, ``lia_program_*``).  ``ProgramRecorder`` has the call surface of
``lia_b200.ops`` that a decode step uses and appends operations to a program instead of launching kernels, so the model
builds its program by running its ordinary decode-step code once against the recorder -- one description of the layer
(decoder.py:172-335 as ``OPTDecoder.layer_rows``), two ways to execute it.  Every operation of a program computes
exactly what the stand-alone kernel computes, in the same order, so both ways give bit-identical results
(tests/test_gpu_decode_program.py).

Replaces the per-step layer loop of lia/modeling_opt.py:1379-1491 (+ models.py:423-431, greedy_search.py:367-395) and
this build's own first-generation CUDA-graph replay of ~340 kernels per step.
"""
import ctypes

import torch

from . import _lib, graphs
from ._lib import EPI_BIAS, EPI_QKV, LiaQkvArgs, check

BF16 = torch.bfloat16


class ProgramUnsupported(_lib.LiaError):
    """The step contains something a program cannot express (a host-side collective, a streamed layer, too many cached
    positions for the score buffer): the caller keeps the kernel-per-operation path."""


class DecodeProgram:
    def __init__(self, rows):
        self.rows = int(rows)
        self.handle = _lib.load().lia_program_create(self.rows)
        if not self.handle:
            raise ProgramUnsupported(f"lia_program_create: {_lib.last_error()}")
        self.keep = []            # tensors the program reads or writes: they must outlive it
        self.finalized = False

    def _p(self, t, name, dtype=BF16, allow_none=False):
        if t is None:
            if allow_none:
                return None
            raise _lib.LiaError(f"{name}: missing tensor")
        if not (t.is_cuda and t.dtype == dtype and t.is_contiguous()):
            raise _lib.LiaError(f"{name}: need a contiguous CUDA {dtype} tensor, got {t.dtype} on {t.device}")
        self.keep.append(t)
        return t.data_ptr()

    def _check(self, rc, what):
        if rc != 0:
            raise ProgramUnsupported(f"{what} failed (code {rc}): {_lib.last_error()}")

    def add_layernorm(self, x, w, b, y, eps):
        h = x.shape[-1]
        self._check(_lib.load().lia_program_add_layernorm(self.handle, self._p(x, "x"), self._p(w, "w"), self._p(b, "b"),
                                                          self._p(y, "y"), x.numel() // h, h, float(eps)), "lia_program_add_layernorm")

    def add_gemm(self, a, w, bias, residual, out, epilogue, qkv=None, tp_args=None):
        M, K = a.shape
        N = w.shape[0]
        if qkv is not None:
            self.keep.append(qkv)
        if tp_args is not None:
            self.keep.append(tp_args)
        self._check(_lib.load().lia_program_add_gemm(
            self.handle, self._p(a, "a"), self._p(w, "w"), self._p(bias, "bias", allow_none=True),
            self._p(residual, "residual", allow_none=True), self._p(out, "out", allow_none=True), M, N, K, int(epilogue),
            ctypes.byref(qkv) if qkv is not None else None, ctypes.byref(tp_args) if tp_args is not None else None), "lia_program_add_gemm")

    def add_attn_decode(self, q, k_cache, v_cache, out, B, b0):
        rows, Bc, H, d = k_cache.shape
        self._check(_lib.load().lia_program_add_attn_decode(self.handle, self._p(q, "q"), self._p(k_cache, "k_cache"),
                                                            self._p(v_cache, "v_cache"), self._p(out, "out"), B, H, d, Bc, b0, rows),
                    "lia_program_add_attn_decode")

    def add_embed(self, embed_tokens, embed_positions, out, B, attention_mask=None):
        t = embed_tokens if embed_tokens is not None else embed_positions
        h = t.shape[1]
        mask_ptr, mask_ld = None, 0
        if attention_mask is not None:
            am = attention_mask
            if not (am.is_cuda and am.dtype == torch.int64 and am.dim() == 2 and am.shape[0] == B and am.stride(1) == 1):
                raise _lib.LiaError("attention_mask: need a CUDA int64 [B, T] tensor with unit column stride")
            self.keep.append(am)
            mask_ptr, mask_ld = am.data_ptr(), am.stride(0) if B > 1 else max(am.stride(0), am.shape[1])
        self._check(_lib.load().lia_program_add_embed(
            self.handle, mask_ptr, mask_ld, self._p(embed_tokens, "embed_tokens", allow_none=True),
            self._p(embed_positions, "embed_positions", allow_none=True), self._p(out, "out"), B, h,
            embed_tokens.shape[0] if embed_tokens is not None else 0, embed_positions.shape[0] if embed_positions is not None else 0),
            "lia_program_add_embed")

    def add_argmax(self, logits):
        B, V = logits.shape
        self._check(_lib.load().lia_program_add_argmax(self.handle, self._p(logits, "logits"), B, V), "lia_program_add_argmax")

    def finalize(self):
        self._check(_lib.load().lia_program_finalize(self.handle), "lia_program_finalize")
        self.finalized = True
        return self

    @property
    def num_ops(self):
        return _lib.load().lia_program_num_ops(self.handle)

    def run(self, pos0, ids_in=None, ids_out=None, suppress_id=-1):
        """One launch on the current stream.  ``pos0`` positions are cached already; ids_in/ids_out are int64 [B] tensors
        (needed when the program holds an embedding / argmax operation)."""
        for t in (ids_in, ids_out):
            if t is not None and not (t.is_cuda and t.dtype == torch.int64 and t.is_contiguous()):
                raise _lib.LiaError("ids: need contiguous CUDA int64 tensors")
        check(_lib.load().lia_program_run(self.handle, int(pos0), ids_in.data_ptr() if ids_in is not None else None,
                                          ids_out.data_ptr() if ids_out is not None else None, int(suppress_id),
                                          torch.cuda.current_stream().cuda_stream), "lia_program_run")
        _lib.launch_count += 1

    def check(self):
        if _lib.load().lia_program_error(self.handle) != 0:
            raise _lib.LiaError(_lib.last_error())

    def close(self):
        if self.handle:
            handle, self.handle = self.handle, None
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            _lib.load().lia_program_destroy(handle)
            self.keep = []

    def __del__(self):
        if getattr(self, "handle", None) and graphs is not None:      # (module globals are gone at interpreter shutdown)
            graphs.finalize(self.close)


class ProgramRecorder:
    """The subset of ``lia_b200.ops`` a decode step calls, recording into a DecodeProgram.  Position-dependent arguments
    (``past_len``, ``pos0``, ``T``, the token ids, the suppressed id) are ignored here: they are arguments of each run."""

    EPI_BIAS, EPI_QKV = EPI_BIAS, EPI_QKV

    def __init__(self, program, device):
        self.program, self.device = program, device

    def layernorm(self, x, w, b, eps=1e-5, out=None):
        if out is None:
            out = torch.empty_like(x)
        self.program.add_layernorm(x, w, b, out, eps)
        return out

    def qkv_args(self, q_out, k_cache, v_cache, S, pos0, b0, scale):
        if S != 1:
            raise ProgramUnsupported("programs append one position per run")
        hq = k_cache.shape[2] * k_cache.shape[3]
        self.program.keep += [q_out, k_cache, v_cache]
        return LiaQkvArgs(q_out.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(), hq, 1, 0, k_cache.shape[1], b0, float(scale))

    def gemm(self, a, w, bias, out=None, epilogue=EPI_BIAS, residual=None, qkv=None, workspace=None):
        if epilogue != EPI_QKV and out is None:
            out = torch.empty(a.shape[0], w.shape[0], dtype=BF16, device=a.device)
        self.program.add_gemm(a, w, bias, residual, out, epilogue, qkv=qkv)
        return out

    def gemm_allreduce(self, a, w, bias, residual, out, tp_args, workspace=None):
        self.program.add_gemm(a, w, bias, residual, out, 0, tp_args=tp_args)
        return out

    def attn_decode(self, q, k_cache, v_cache, B, T, b0=0, out=None, splits=0, workspace=None):
        if out is None:
            out = torch.empty(B, k_cache.shape[2] * k_cache.shape[3], dtype=BF16, device=q.device)
        self.program.add_attn_decode(q, k_cache, v_cache, out, B, b0)
        return out

    def embed(self, ids, embed_tokens, embed_positions, past_len, out=None, attention_mask=None):
        B, S = ids.shape
        if S != 1:
            raise ProgramUnsupported("programs embed one position per run")
        t = embed_tokens if embed_tokens is not None else embed_positions
        if out is None:
            out = torch.empty(B, 1, t.shape[1], dtype=BF16, device=t.device)
        self.program.add_embed(embed_tokens, embed_positions, out, B, attention_mask)
        return out

    def argmax(self, logits, suppress_id=-1, out=None):
        self.program.add_argmax(logits)
        return out

    def _unsupported(self, name):
        def f(*a, **k):
            raise ProgramUnsupported(f"ops.{name} cannot be part of a decode program")
        return f

    def __getattr__(self, name):
        if name in ("attn_prefill", "kv_append", "residual_add"):
            return self._unsupported(name)
        raise AttributeError(name)
