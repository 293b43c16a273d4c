"""The "cuda" table of the reference's operator registry.

The reference's op-level plugin API is ``ipex.llm.modules``
(intel_extension_for_pytorch/llm/modules/__init__.py:1-19): a device-type-keyed registry
``IPEXRuntimeCustomOps.fusion_modules = {"cpu": CPU_fusion_modules}`` indexed by the
``IPEXCustomOpType`` enum (llm/modules/utils.py:24-93), which asserts ``device_type in ["cpu"]``
(:79-82).  This module supplies the missing ``"cuda"`` entries for the four op types on the OPT
path, with the same constructor / call signatures, each backed by one C-ABI entry point:

  LINEAR_RELU             LinearRelu(linear)(x)        = relu(linear(x))     linear_fusion.py:112-138
  LINEAR_ADD              LinearAdd(linear)(x, y)      = linear(x) + y       linear_fusion.py:256-283
  FAST_LAYERNORM          FastLayerNorm(shape, eps, w, b)(h) / .apply(...)   mha_fusion.py:148-215
  INDIRECTACCESS_KVCACHE  IndirectAccessKVCache(text_max_length).apply(q, k, v, scale_attn,
                          layer_past, head_mask, attention_mask)             mha_fusion.py:503-620
"""
import math
from enum import Enum

import torch

from . import ops
from .ops import EPI_BIAS_RELU, EPI_BIAS_RESIDUAL

BF16 = torch.bfloat16
DEVICE = "cuda"        # where the ops keep their parameters (the CPU suite points this at its kernel stand-in's device)


class IPEXCustomOpType(Enum):          # values as llm/modules/utils.py:24-40
    LINEAR_RELU = 3
    LINEAR_ADD = 6
    FAST_LAYERNORM = 12
    INDIRECTACCESS_KVCACHE = 14


class _LinearFusionCUDA:
    """Holds the Linear's weight [N,K] / bias [N] as contiguous CUDA bf16 (plain row-major; no
    TPP blocking, cf. nn/utils/_weight_prepack.py:19-63)."""

    def __init__(self, linear):
        self.weight = linear.weight.detach().to(DEVICE, BF16).contiguous()
        self.bias = None if linear.bias is None else linear.bias.detach().to(DEVICE, BF16).contiguous()
        n, k = self.weight.shape
        self._ws = {}

    def _workspace(self, m):
        if m not in self._ws:
            n, k = self.weight.shape
            self._ws[m] = ops.GemmWorkspace(ops.GemmWorkspace.bytes_for([(m, n, k)]), self.weight.device)
        return self._ws[m]

    def _rows(self, x):
        return x.reshape(-1, x.shape[-1]).contiguous()


class LinearReluCUDA(_LinearFusionCUDA):
    def __call__(self, x):
        r = self._rows(x)
        y = ops.gemm(r, self.weight, self.bias, epilogue=EPI_BIAS_RELU, workspace=self._workspace(r.shape[0]))
        return y.view(*x.shape[:-1], self.weight.shape[0])


class LinearAddCUDA(_LinearFusionCUDA):
    def __call__(self, x, y):
        r = self._rows(x)
        o = ops.gemm(r, self.weight, self.bias, epilogue=EPI_BIAS_RESIDUAL, residual=self._rows(y),
                     workspace=self._workspace(r.shape[0]))
        return o.view(*x.shape[:-1], self.weight.shape[0])


class FastLayerNormCUDA:
    def __init__(self, normalized_shape, eps, weight, bias=None):
        self.normalized_shape, self.eps = normalized_shape, eps
        self.weight = weight.detach().to(DEVICE, BF16).contiguous()
        self.bias = (torch.zeros_like(self.weight) if bias is None else bias.detach().to(DEVICE, BF16).contiguous())

    def __call__(self, hidden_states):
        return ops.layernorm(hidden_states.contiguous(), self.weight, self.bias, self.eps)

    @classmethod
    def apply(cls, hidden_states, normalized_shape, weight, bias, eps):
        return ops.layernorm(hidden_states.contiguous(), weight, bias if bias is not None else torch.zeros_like(weight), eps)


class IndirectAccessKVCacheCUDA:
    """layer_past = (seq_info [1,1,T,T] (only its shape is read), key_cache [max_seq,B,H,d],
    value_cache [max_seq,B,H,d], beam_idx [max_seq,B]); first call passes None and the caches are
    allocated for ``text_max_length`` positions (MaskedMultiHeadAttentionKrnl.cpp:1384-1392)."""

    def __init__(self, text_max_length=2048):
        self.text_max_length = text_max_length

    @classmethod
    def apply(cls, query, key, value, scale_attn, layer_past=None, head_mask=None, attention_mask=None, alibi=None,
              add_casual_mask=True, seq_info=None, text_max_length=0):
        if head_mask is not None or alibi is not None:
            raise NotImplementedError("head_mask / alibi are not supported by the CUDA kernels")
        B, S, H, d = query.shape
        past = 0 if layer_past is None else int(layer_past[0].shape[2])
        if layer_past is None or past == 0:
            tmax = max(int(text_max_length or 0), S + 1) if text_max_length else 2048
            kc = torch.zeros(tmax, B, H, d, dtype=BF16, device=query.device)
            vc = torch.zeros_like(kc)
            beam = torch.zeros(tmax, B, dtype=torch.long, device=query.device)
        else:
            kc, vc, beam = layer_past[1], layer_past[2], layer_past[3]
        q = ops.kv_append(query.contiguous(), key.contiguous(), value.contiguous(), kc, vc, past, 0, 1.0 / float(scale_attn))
        if S != 1:
            if past != 0:
                raise NotImplementedError("multi-token call with a non-empty cache")
            out = ops.attn_prefill(q.view(B * S, H * d), kc, vc, B, S, 0)
        else:
            out = ops.attn_decode(q.view(B, H * d), kc, vc, B, past + 1, 0)
        T = past + S
        new_past = (torch.empty(1, 1, T, T, dtype=torch.long, device="meta"), kc, vc, beam)
        return out.view(B, S, H, d), None, new_past

    def __call__(self, query, key, value, scale_attn, layer_past=None, head_mask=None, attention_mask=None, alibi=None,
                 add_casual_mask=True, seq_info=None):
        return self.apply(query, key, value, scale_attn, layer_past, head_mask, attention_mask, alibi, add_casual_mask,
                          seq_info, self.text_max_length)


CUDA_fusion_modules = {
    IPEXCustomOpType.LINEAR_RELU: LinearReluCUDA,
    IPEXCustomOpType.LINEAR_ADD: LinearAddCUDA,
    IPEXCustomOpType.FAST_LAYERNORM: FastLayerNormCUDA,
    IPEXCustomOpType.INDIRECTACCESS_KVCACHE: IndirectAccessKVCacheCUDA,
}

# what a maintainer adds to IPEXRuntimeCustomOps.fusion_modules (llm/modules/utils.py:66-68)
fusion_modules = {"cuda": CUDA_fusion_modules}

# user-facing names, as exported by ipex.llm.modules (llm/modules/__init__.py)
LinearRelu, LinearAdd, FastLayerNorm, IndirectAccessKVCache = (LinearReluCUDA, LinearAddCUDA, FastLayerNormCUDA,
                                                               IndirectAccessKVCacheCUDA)
