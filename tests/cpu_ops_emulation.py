"""TEST INFRASTRUCTURE, NOT PRODUCT: a stock-PyTorch CPU stand-in for ``lia_b200.ops`` so that the HOST logic of
``modeling_opt`` (layer order, minibatch loop, cache bookkeeping, the greedy loop, the forward/generate faces) can run in
the CPU suite (``-m "not gpu"``) and be compared with the oracle.

Only ``tests/`` may import this module, through the ``cpu_ops`` fixture below; it monkeypatches the ``ops`` names the
model calls and the three ``torch.cuda`` entry points ``generate()`` touches, and undoes all of it when the test ends.
The product never falls back to it: ``lia_b200.ops`` itself still rejects non-CUDA tensors (tests/test_c_abi.py).

Every stand-in computes what the kernel it replaces is specified to compute (include/lia_b200.h), with the reference's
rounding points, using the same PyTorch ops as ``oracle/opt_ref.py`` -- so on one CPU thread the model
built on it must agree with the oracle bit for bit; any difference is a host-logic bug (wrong buffer, order, offset).
"""
import pytest
import torch
from torch.nn import functional as F

BF16 = torch.bfloat16
EPI_BIAS, EPI_BIAS_RELU, EPI_BIAS_RESIDUAL, EPI_QKV = 0, 1, 2, 3
MASK_VALUE = -3.4028e+38


class QkvArgs:
    """Stand-in for LiaQkvArgs: holds the tensors instead of their addresses."""

    def __init__(self, q_out, k_cache, v_cache, S, pos0, b0, scale):
        self.q_out, self.k_cache, self.v_cache, self.S, self.pos0, self.b0, self.scale = q_out, k_cache, v_cache, S, pos0, b0, scale


class GemmWorkspace:
    def __init__(self, nbytes, device):
        self.buf = torch.zeros(16, dtype=torch.uint8, device=device)

    @staticmethod
    def bytes_for(shapes):
        return 16


class Calls:
    """What the model asked for, in order: tests assert on launch sequences."""

    def __init__(self):
        self.log = []

    def names(self):
        return [c[0] for c in self.log]


def make(calls):
    def layernorm(x, w, b, eps=1e-5, out=None):
        calls.log.append(("layernorm", tuple(x.shape)))
        y = F.layer_norm(x, (x.shape[-1],), w, b, eps)
        if out is None:
            return y
        out.copy_(y.view(out.shape))
        return out

    def _mm_bias(a, w, bias):
        r = torch.matmul(a, w.t())                        # r1 = bf16(acc)
        return r + bias if bias is not None else r        # r2 = bf16(r1 + bias)

    def gemm(a, w, bias, out=None, epilogue=EPI_BIAS, residual=None, qkv=None, workspace=None):
        calls.log.append(("gemm", epilogue, a.shape[0], w.shape[0], a.shape[1]))
        assert a.dtype == BF16 and w.dtype == BF16 and a.is_contiguous() and w.is_contiguous()
        if epilogue == EPI_QKV:
            hq = qkv.k_cache.shape[2] * qkv.k_cache.shape[3]
            assert w.shape[0] == 3 * hq
            M = a.shape[0]
            nb, S = M // qkv.S, qkv.S
            # three products, as the reference issues them (attentions.py:393-394, 418)
            k = _mm_bias(a, w[hq:2 * hq], bias[hq:2 * hq])
            v = _mm_bias(a, w[2 * hq:], bias[2 * hq:])
            q = _mm_bias(a, w[:hq], bias[:hq])
            qkv.q_out[:M].copy_(q * qkv.scale)
            H, d = qkv.k_cache.shape[2], qkv.k_cache.shape[3]
            qkv.k_cache[qkv.pos0:qkv.pos0 + S, qkv.b0:qkv.b0 + nb].copy_(k.view(nb, S, H, d).permute(1, 0, 2, 3))
            qkv.v_cache[qkv.pos0:qkv.pos0 + S, qkv.b0:qkv.b0 + nb].copy_(v.view(nb, S, H, d).permute(1, 0, 2, 3))
            return None
        r = _mm_bias(a, w, bias)
        if epilogue == EPI_BIAS_RELU:
            r = F.relu(r)
        elif epilogue == EPI_BIAS_RESIDUAL:
            assert residual is not None and residual.shape == r.shape
            r = residual + r
        if out is None:
            return r
        assert out.shape == r.shape and out.is_contiguous()
        out.copy_(r)
        return out

    def qkv_args(q_out, k_cache, v_cache, S, pos0, b0, scale):
        return QkvArgs(q_out, k_cache, v_cache, S, pos0, b0, scale)

    def kv_append(q, k, v, k_cache, v_cache, pos0, b0, scale, q_out=None):
        calls.log.append(("kv_append", tuple(q.shape), pos0, b0))
        B, S = q.shape[0], q.shape[1]
        k_cache[pos0:pos0 + S, b0:b0 + B].copy_(k.permute(1, 0, 2, 3))
        v_cache[pos0:pos0 + S, b0:b0 + B].copy_(v.permute(1, 0, 2, 3))
        r = q * scale
        if q_out is None:
            return r
        q_out.copy_(r)
        return q_out

    def _attention(q, k_cache, v_cache, B, S, T, b0, causal):
        _, Bc, H, d = k_cache.shape
        qh = q[:B * S].view(B, S, H, d).transpose(1, 2).contiguous().view(B * H, S, d)
        k = k_cache[:T, b0:b0 + B].permute(1, 2, 0, 3).contiguous().view(B * H, T, d)
        v = v_cache[:T, b0:b0 + B].permute(1, 2, 0, 3).contiguous().view(B * H, T, d)
        attn = torch.bmm(qh, k.transpose(1, 2))
        if causal:
            mask = torch.triu(torch.ones(S, T), diagonal=1) * MASK_VALUE
            attn = attn.view(B, H, S, T) + mask
            attn = torch.max(attn, torch.tensor(torch.finfo(attn.dtype).min))
        attn = torch.softmax(attn.view(B * H, S, T), dim=-1, dtype=BF16)
        ctx = torch.bmm(attn, v)
        return ctx.view(B, H, S, d).transpose(1, 2).reshape(B * S, H * d)

    def attn_prefill(q, k_cache, v_cache, B, S, b0=0, out=None):
        calls.log.append(("attn_prefill", B, S, b0))
        ctx = _attention(q, k_cache, v_cache, B, S, S, b0, True)
        if out is None:
            return ctx
        out[:B * S].copy_(ctx)
        return out

    def attn_decode(q, k_cache, v_cache, B, T, b0=0, out=None, splits=0, workspace=None):
        calls.log.append(("attn_decode", B, T, b0))
        ctx = _attention(q, k_cache, v_cache, B, 1, T, b0, False)
        if out is None:
            return ctx
        out[:B].copy_(ctx)
        return out

    def attn_decode_workspace(B, H, d, device, max_splits=32):
        return torch.empty(1, dtype=torch.float32, device=device)

    def embed(ids, embed_tokens, embed_positions, past_len, out=None, attention_mask=None):
        calls.log.append(("embed", tuple(ids.shape), past_len))
        B, S = ids.shape
        # the C ABI's argument checks (csrc/misc.cu lia_embed_masked_bf16), so that contract violations fail on CPU too
        assert embed_tokens is not None or embed_positions is not None
        assert past_len >= 0 and (embed_positions is None or past_len + S + 2 <= embed_positions.shape[0]), \
            "positions exceed the table"
        assert attention_mask is None or attention_mask.shape[1] >= past_len + S
        V = embed_tokens.shape[0] if embed_tokens is not None else 0
        P = embed_positions.shape[0] if embed_positions is not None else 0
        if attention_mask is None:
            pos = torch.arange(past_len, past_len + S).expand(B, S)
        else:
            am = attention_mask[:, :past_len + S].long()
            pos = ((torch.cumsum(am, dim=1) * am) - 1)[:, past_len:]
        if embed_positions is None:                       # a null table contributes nothing (rows copied through)
            r = F.embedding(ids.clamp(0, V - 1), embed_tokens)
        elif embed_tokens is None:
            r = F.embedding((pos + 2).clamp(0, P - 1), embed_positions)
        else:
            pos = (pos + 2).clamp(0, P - 1)               # the kernel clamps both lookups
            r = F.embedding(ids.clamp(0, V - 1), embed_tokens) + F.embedding(pos, embed_positions)
        if out is None:
            return r
        out.copy_(r.view(out.shape))
        return out

    def argmax(logits, suppress_id=-1, out=None):
        calls.log.append(("argmax", suppress_id))
        lg = logits.float().clone()
        if suppress_id >= 0:
            lg[:, suppress_id] = -float("inf")
        r = torch.argmax(lg, dim=-1)
        if out is None:
            return r
        out.copy_(r)
        return out

    def residual_add(x, residual, out=None):
        calls.log.append(("residual_add", tuple(x.shape)))
        r = residual + x
        if out is None:
            return r
        out.copy_(r)
        return out

    def gemm_allreduce(*a, **k):
        raise AssertionError("the fused NVLink exchange has no CPU stand-in (LIA_TP_FUSED=0 path only)")

    return dict(layernorm=layernorm, gemm=gemm, qkv_args=qkv_args, kv_append=kv_append, attn_prefill=attn_prefill, attn_decode=attn_decode,
                attn_decode_workspace=attn_decode_workspace, embed=embed, argmax=argmax, residual_add=residual_add,
                gemm_allreduce=gemm_allreduce, GemmWorkspace=GemmWorkspace)


class _Event:
    def __init__(self, enable_timing=False):
        pass

    def record(self, *a):
        pass

    def elapsed_time(self, other):
        return 0.0


@pytest.fixture
def cpu_ops(monkeypatch):
    """Patch lia_b200.ops (and the torch.cuda calls generate() makes) for the duration of one test; yields the call log."""
    import lia_b200  # noqa: F401
    from lia_b200 import ops
    calls = Calls()
    for name, fn in make(calls).items():
        monkeypatch.setattr(ops, name, fn)
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda *a, **k: None)
    torch.set_num_threads(1)
    yield calls
