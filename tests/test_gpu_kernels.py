"""Kernel-level parity on a real B200: every C-ABI entry point against the same op restated
in plain PyTorch (fp32 math with the reference's bf16 rounding points)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

BF16 = torch.bfloat16


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import lia_b200  # noqa: F401
    from lia_b200 import ops as o
    return o


def rnd(*shape, std=1.0, seed=0, device="cuda"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * std).to(BF16).to(device)


def r16(x):
    return x.to(BF16).float()


def close_bf16(got, ref, ulps=2.0, frac_exact=0.0, what="", scale=None):
    """|got-ref| <= ulps bf16 ulps (2^-7 relative) of ``scale`` -- the magnitude of the largest
    intermediate that was rounded on the way to each element (default: |ref| itself)."""
    got, ref = got.float(), ref.float()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert torch.isfinite(got).all(), what
    mag = ref.abs() if scale is None else torch.maximum(ref.abs(), scale.float().abs().expand_as(ref))
    tol = ulps * (2.0 ** -7) * mag.clamp_min(ref.abs().max() * 2e-2 + 1e-30)
    bad = (got - ref).abs() > tol
    assert not bad.any(), f"{what}: {int(bad.sum())} / {bad.numel()} beyond {ulps} ulp; max abs err {(got - ref).abs().max().item():.4g}"
    if frac_exact:
        assert (got == ref).float().mean().item() >= frac_exact, what


# ------------------------------------------------------------------ layernorm
@pytest.mark.parametrize("rows,h", [(1, 64), (5, 128), (300, 2048), (64, 7168), (17, 12288), (3, 192)])
def test_layernorm(ops, rows, h):
    x = rnd(rows, h, seed=1) * 3 + 0.5
    w = rnd(h, seed=2) * 0.1 + 1
    b = rnd(h, seed=3) * 0.1
    y = ops.layernorm(x, w, b)
    ref = torch.nn.functional.layer_norm(x.float(), (h,), w.float(), b.float(), 1e-5)
    close_bf16(y, r16(ref), ulps=1.01, frac_exact=0.9, what=f"ln {rows}x{h}")


# ------------------------------------------------------------------ gemm
def gemm_ref(a, w, bias, epilogue, residual=None, with_scale=False):
    acc = a.float() @ w.float().t()
    r = r16(acc)
    scale = r.abs()
    if bias is not None:
        r = r16(r + bias.float())
        scale = torch.maximum(scale, r.abs())
    if epilogue == 1:
        r = torch.relu(r)
    elif epilogue == 2:
        r = r16(residual.float() + r)
    return (r, scale) if with_scale else r


SWAP_SHAPES = [(1, 128, 128), (3, 192, 200), (16, 256, 64), (17, 384, 1024), (64, 7168, 7168), (64, 21504, 7168),
               (100, 768, 512), (128, 1024, 4096), (64, 7168, 28672), (8, 50272, 2048)]
NORMAL_SHAPES = [(129, 128, 64), (256, 256, 128), (300, 392, 200), (1024, 768, 512), (2048, 7168, 7168),
                 (4096, 28672, 7168), (2048, 7168, 28672), (200, 50272, 256)]


@pytest.mark.parametrize("M,N,K", SWAP_SHAPES + NORMAL_SHAPES)
@pytest.mark.parametrize("epilogue", [0, 1, 2])
def test_gemm(ops, M, N, K, epilogue):
    if epilogue != 0 and M * N > 4096 * 28672 // 2 and epilogue == 1:
        pass
    a = rnd(M, K, seed=M + N)
    w = rnd(N, K, std=K ** -0.5, seed=K + 1)
    bias = rnd(N, std=0.5, seed=5)
    res = rnd(M, N, seed=6) if epilogue == 2 else None
    ws = ops.GemmWorkspace(ops.GemmWorkspace.bytes_for([(M, N, K)]), "cuda")
    for rep in range(2):   # second pass checks that the split-K counters were left zeroed
        y = ops.gemm(a, w, bias, epilogue=epilogue, residual=res, workspace=ws)
        ref, scale = gemm_ref(a, w, bias, epilogue, res, with_scale=True)
        close_bf16(y, ref, ulps=2.01, what=f"gemm {M}x{N}x{K} epi {epilogue} rep {rep}", scale=scale)
    # without a workspace (no split-K) and without bias
    y2 = ops.gemm(a, w, None, epilogue=0)
    close_bf16(y2, gemm_ref(a, w, None, 0), ulps=1.01, what=f"gemm nobias {M}x{N}x{K}")


def test_gemm_deterministic(ops):
    a, w, bias = rnd(64, 7168, seed=1), rnd(7168, 7168, std=0.01, seed=2), rnd(7168, seed=3)
    ws = ops.GemmWorkspace(ops.GemmWorkspace.bytes_for([(64, 7168, 7168)]), "cuda")
    y0 = ops.gemm(a, w, bias, workspace=ws).clone()
    for _ in range(5):
        assert torch.equal(ops.gemm(a, w, bias, workspace=ws), y0)


@pytest.mark.parametrize("B,S,H,d,pos0,b0,Bc", [(3, 8, 2, 64, 0, 0, 3), (2, 5, 1, 128, 0, 1, 4), (64, 1, 8, 128, 17, 0, 64),
                                               (4, 256, 8, 128, 0, 4, 8), (5, 1, 4, 64, 3, 2, 9)])
def test_gemm_qkv_scatter(ops, B, S, H, d, pos0, b0, Bc):
    hq, K = H * d, 256
    M = B * S
    a = rnd(M, K, seed=1)
    w = rnd(3 * hq, K, std=K ** -0.5, seed=2)
    bias = rnd(3 * hq, std=0.5, seed=3)
    Tmax = pos0 + S + 2
    kc = torch.full((Tmax, Bc, H, d), 7.0, dtype=BF16, device="cuda")
    vc = torch.full((Tmax, Bc, H, d), 7.0, dtype=BF16, device="cuda")
    q = torch.empty(M, hq, dtype=BF16, device="cuda")
    scale = d ** -0.5
    ws = ops.GemmWorkspace(ops.GemmWorkspace.bytes_for([(M, 3 * hq, K)]), "cuda")
    ops.gemm(a, w, bias, epilogue=ops.EPI_QKV, qkv=ops.qkv_args(q, kc, vc, S, pos0, b0, scale), workspace=ws)
    r2, mag = gemm_ref(a, w, bias, 0, with_scale=True)
    close_bf16(q, r16(r2[:, :hq] * torch.tensor(scale, dtype=torch.float32)), ulps=2.01, what="q", scale=mag[:, :hq] * scale)
    k_ref = r2[:, hq:2 * hq].view(B, S, H, d).permute(1, 0, 2, 3)
    v_ref = r2[:, 2 * hq:].view(B, S, H, d).permute(1, 0, 2, 3)
    k_mag = mag[:, hq:2 * hq].view(B, S, H, d).permute(1, 0, 2, 3)
    v_mag = mag[:, 2 * hq:].view(B, S, H, d).permute(1, 0, 2, 3)
    close_bf16(kc[pos0:pos0 + S, b0:b0 + B], k_ref, ulps=2.01, what="k", scale=k_mag)
    close_bf16(vc[pos0:pos0 + S, b0:b0 + B], v_ref, ulps=2.01, what="v", scale=v_mag)
    # everything outside the window untouched
    mask = torch.ones_like(kc, dtype=torch.bool)
    mask[pos0:pos0 + S, b0:b0 + B] = False
    assert (kc[mask] == 7.0).all() and (vc[mask] == 7.0).all()


# ------------------------------------------------------------------ attention
def attn_ref(q, k, v, causal):
    """q [B,H,S,d] (scaled), k/v [B,H,T,d]: the reference's bmm / softmax(dtype=bf16) / bmm rounding points."""
    s = r16(q.float() @ k.float().transpose(-1, -2))
    if causal:
        S, T = s.shape[-2:]
        s = s.masked_fill(torch.triu(torch.ones(S, T, dtype=torch.bool, device=s.device), 1), float("-inf"))
    p = r16(torch.softmax(s, dim=-1))
    return r16(p @ v.float())


@pytest.mark.parametrize("B,S,H,d,b0,Bc", [(2, 5, 1, 128, 0, 2), (3, 8, 2, 64, 1, 5), (2, 64, 3, 128, 0, 2), (2, 65, 2, 64, 0, 2),
                                          (2, 256, 4, 128, 2, 4), (1, 200, 2, 64, 0, 1), (1, 512, 2, 128, 0, 1)])
def test_attn_prefill(ops, B, S, H, d, b0, Bc):
    q = rnd(B, S, H, d, std=0.3, seed=1)
    kc = rnd(S + 3, Bc, H, d, seed=2)
    vc = rnd(S + 3, Bc, H, d, seed=3)
    out = ops.attn_prefill(q.view(B * S, H * d), kc, vc, B, S, b0)
    ref = attn_ref(q.permute(0, 2, 1, 3), kc[:S, b0:b0 + B].permute(1, 2, 0, 3), vc[:S, b0:b0 + B].permute(1, 2, 0, 3), True)
    ref = ref.permute(0, 2, 1, 3).reshape(B * S, H * d)
    close_bf16(out, ref, ulps=3.01, what=f"prefill attn S={S} d={d}", scale=ref.abs().amax(-1, keepdim=True))


@pytest.mark.parametrize("B,T,H,d,b0,Bc,splits", [(2, 1, 1, 128, 0, 2, 1), (3, 9, 2, 64, 1, 5, 1), (64, 288, 8, 128, 0, 64, 1),
                                                 (2, 1000, 4, 128, 0, 2, 0), (1, 2048, 2, 64, 0, 1, 0), (4, 300, 4, 128, 4, 8, 3)])
def test_attn_decode(ops, B, T, H, d, b0, Bc, splits):
    q = rnd(B, H, d, std=0.3, seed=1)
    kc = rnd(T + 2, Bc, H, d, seed=2)
    vc = rnd(T + 2, Bc, H, d, seed=3)
    ws = ops.attn_decode_workspace(B, H, d, "cuda")
    out = ops.attn_decode(q.view(B, H * d), kc, vc, B, T, b0, splits=splits, workspace=ws)
    ref = attn_ref(q.view(B, H, 1, d), kc[:T, b0:b0 + B].permute(1, 2, 0, 3), vc[:T, b0:b0 + B].permute(1, 2, 0, 3), False)
    ref = ref.reshape(B, H * d)
    close_bf16(out, ref, ulps=2.01 if splits == 1 else 4.01, what=f"decode attn T={T} d={d} splits={splits}",
               scale=ref.abs().amax(-1, keepdim=True))


# ------------------------------------------------------------------ small kernels
def test_embed_argmax_residual(ops):
    V, h, B, S = 1000, 256, 3, 7
    tok, pos = rnd(V, h, seed=1), rnd(64, h, seed=2)
    ids = torch.randint(0, V, (B, S), device="cuda")
    out = ops.embed(ids, tok, pos, 5)
    ref = r16(tok[ids].float() + pos[torch.arange(S, device="cuda") + 5 + 2][None].float())
    assert torch.equal(out.float(), ref)
    logits = rnd(B, 50272, seed=3)
    logits[0, 2] = 100.0
    logits[1, 10] = logits[1, 20] = 50.0       # tie -> lowest index
    nxt = ops.argmax(logits, suppress_id=2)
    lf = logits.float().clone()
    lf[:, 2] = float("-inf")
    assert torch.equal(nxt, lf.argmax(-1))
    assert nxt[1].item() == 10
    assert torch.equal(ops.argmax(logits, -1), logits.float().argmax(-1))
    a, b = rnd(33, 256, seed=4), rnd(33, 256, seed=5)
    assert torch.equal(ops.residual_add(a, b).float(), r16(b.float() + a.float()))


def test_embed_positions_from_padded_mask_vs_reference_golden(ops, golden_dir):
    """lia_embed_masked_bf16 against rows produced by the reference's own OPTLearnedPositionalEmbedding.forward
    (tests/golden/positions_padded.npz): left/right padding, a hole, a fully padded row; prefill and decode steps."""
    import os
    import numpy as np
    z = np.load(os.path.join(golden_dir, "positions_padded.npz"))
    bits = lambda a: torch.from_numpy(a.view(np.int16).copy()).view(BF16)   # noqa: E731
    table = bits(z["table"]).cuda()
    B, S, new, h = int(z["B"]), int(z["S"]), int(z["new"]), int(z["h"])
    tok = torch.zeros(8, h, dtype=BF16, device="cuda")                      # token rows are zero: out == position row
    buf = torch.ones(B, S + new + 3, dtype=torch.int64, device="cuda")      # wider than needed: row stride != columns
    buf[:, :S] = torch.from_numpy(z["mask"]).cuda()
    for step in range(new + 1):
        past, s = (0, S) if step == 0 else (S + step - 1, 1)
        ids = torch.zeros(B, s, dtype=torch.int64, device="cuda")
        out = ops.embed(ids, tok, table, past, attention_mask=buf[:, :past + s])
        assert torch.equal(out.view(torch.int16).cpu(), bits(z[f"rows{step}"]).view(torch.int16)), step
    # no mask == all-ones mask
    ids = torch.randint(0, 8, (B, S), device="cuda")
    tok = rnd(8, h, seed=9)
    assert torch.equal(ops.embed(ids, tok, table, 3), ops.embed(ids, tok, table, 3, attention_mask=torch.ones(B, S + 3, dtype=torch.int64, device="cuda")))
    from lia_b200._lib import LiaError
    with pytest.raises(LiaError):
        ops.embed(ids, tok, table, 3, attention_mask=torch.ones(B, S, dtype=torch.int64, device="cuda"))   # too few columns


def test_errors_are_loud(ops):
    from lia_b200._lib import LiaError
    a = rnd(4, 12)          # K not a multiple of 8
    w = rnd(16, 12)
    with pytest.raises(LiaError):
        ops.gemm(a, w, None)
    with pytest.raises(LiaError):
        ops.layernorm(torch.zeros(2, 8), torch.zeros(8), torch.zeros(8))     # CPU tensors: no fallback
