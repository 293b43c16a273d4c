"""lia_b200 -- B200-native (sm_100a) implementation of LIA's OPT decoder-layer hot path.

Layout:
  csrc/            CUDA kernels + the C ABI (include/lia_b200.h) -> libliab200.so
  _lib.py          ctypes loader (fails loudly when the library is missing)
  ops.py           tensor-level wrappers over the C ABI
  llm_modules.py   "cuda" table of the reference's operator registry (ipex.llm.modules)
  modeling_opt.py  OPTForCausalLM / OPTDecoder / OPTDecoderLayer with the LIA kwargs
  weights.py       per-layer weight slabs, random-init and dummy-weight generators
  streamer.py      pinned-host layer streaming (replaces AMX-CPU compute + CXL tiering)
  kv_spill.py      per-layer KV-cache spill to pinned host memory (K/V larger than HBM)
  checkpoint.py    HF safetensors/.bin reader and the native per-layer slab format
  tp.py            tensor-parallel plumbing: peer arenas for the fused projection+all-reduce kernel, NCCL bootstrap
  run.py           the reference's run.py / run_generation.py command line
"""
from . import _lib  # noqa: F401
from .modeling_opt import OPTConfig, OPTForCausalLM, OPT_CONFIGS  # noqa: F401

__all__ = ["OPTConfig", "OPTForCausalLM", "OPT_CONFIGS"]
