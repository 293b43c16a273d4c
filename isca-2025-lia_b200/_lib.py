"""ctypes binding of libliab200.so (C ABI declared in include/lia_b200.h).

The reference's only C-ABI precedent is loaded the same way
(lia/cxl/numa_alloc.py:8-26: ctypes.CDLL + argtypes/restype).  There is NO fallback:
if the shared library is missing or a call fails, this module raises.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libliab200.so")

EPI_BIAS, EPI_BIAS_RELU, EPI_BIAS_RESIDUAL, EPI_QKV = 0, 1, 2, 3
ABI_VERSION = 5
TP_MAX_WORLD = 8
P2P_HANDLE_BYTES = 64


class LiaQkvArgs(ctypes.Structure):
    _fields_ = [("q_out", c_void_p), ("k_cache", c_void_p), ("v_cache", c_void_p),
                ("hq", c_int32), ("S", c_int32), ("pos0", c_int32), ("cache_batch", c_int32),
                ("b0", c_int32), ("q_scale", c_float)]


class LiaTpArgs(ctypes.Structure):
    _fields_ = [("rank", c_int32), ("world", c_int32), ("arena", c_void_p * TP_MAX_WORLD),
                ("ctl_off", ctypes.c_uint64), ("recv_off", ctypes.c_uint64), ("recv_bytes", ctypes.c_uint64),
                ("out_off", ctypes.c_uint64), ("mc_arena", c_void_p)]


class LiaError(RuntimeError):
    pass


_PROTOTYPES = {
    "lia_abi_version": (c_int, []),
    "lia_last_error": (c_char_p, []),
    "lia_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "lia_layernorm_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p]),
    "lia_gemm_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "lia_gemm_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                              POINTER(LiaQkvArgs), c_void_p, c_size_t, c_void_p]),
    "lia_attn_prefill_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                      c_int, c_void_p]),
    "lia_attn_decode_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "lia_attn_decode_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                     c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "lia_kv_append_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                   c_int, c_int, c_float, c_void_p]),
    "lia_embed_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                               c_void_p]),
    "lia_embed_masked_bf16": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                      c_int, c_int, c_void_p]),
    "lia_argmax_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "lia_residual_add_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "lia_tp_ctl_bytes": (c_size_t, []),
    "lia_tp_recv_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "lia_gemm_allreduce_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                        POINTER(LiaTpArgs), c_void_p, c_size_t, c_void_p]),
    "lia_tp_error": (c_int, [POINTER(LiaTpArgs)]),
    "lia_p2p_alloc": (c_int, [c_size_t, POINTER(c_void_p), c_void_p]),
    "lia_p2p_open": (c_int, [c_void_p, POINTER(c_void_p)]),
    "lia_p2p_close": (c_int, [c_void_p]),
    "lia_p2p_free": (c_int, [c_void_p]),
    "lia_host_arena_alloc": (c_void_p, [c_size_t]),
    "lia_host_arena_free": (c_int, [c_void_p, c_size_t]),
    "lia_streamer_create": (c_void_p, [POINTER(c_void_p), c_int, c_size_t]),
    "lia_streamer_prefetch": (c_int, [c_void_p, c_int, c_void_p, c_size_t]),
    "lia_streamer_wait": (c_int, [c_void_p, c_int, c_void_p]),
    "lia_streamer_release": (c_int, [c_void_p, c_int, c_void_p]),
    "lia_streamer_stats": (c_int, [c_void_p, POINTER(c_double), POINTER(c_double)]),
    "lia_streamer_destroy": (c_int, [c_void_p]),
    "lia_program_create": (c_void_p, [c_int]),
    "lia_program_add_layernorm": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float]),
    "lia_program_add_gemm": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                     POINTER(LiaQkvArgs), POINTER(LiaTpArgs)]),
    "lia_program_add_attn_decode": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int]),
    "lia_program_add_embed": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int]),
    "lia_program_add_argmax": (c_int, [c_void_p, c_void_p, c_int, c_int]),
    "lia_program_finalize": (c_int, [c_void_p]),
    "lia_program_run": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "lia_program_error": (c_int, [c_void_p]),
    "lia_program_num_ops": (c_int, [c_void_p]),
    "lia_program_destroy": (c_int, [c_void_p]),
}

EXPORTS = tuple(_PROTOTYPES)

_lib = None


def load():
    """Load libliab200.so once; raise (never fall back) if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LiaError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C isca-2025-lia_b200/csrc` (there is no CPU/PyTorch fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOTYPES.items():
            fn = getattr(lib, name)     # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if lib.lia_abi_version() != ABI_VERSION:
            raise LiaError(f"ABI mismatch: library {lib.lia_abi_version()} vs binding {ABI_VERSION}")
        _lib = lib
    return _lib


def last_error():
    msg = load().lia_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc, what):
    if rc != 0:
        raise LiaError(f"{what} failed (code {rc}): {last_error()}")


launch_count = 0   # kernels launched through this binding (bench.py's gpu_launches)
