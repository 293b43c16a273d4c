"""Parity test of the CTA-pair (tcgen05.mma.cta_group::2) prefill GEMM -- the default for M >= 512 -- against the
one-CTA 128 x 256 kernel (LIA_GEMM_2CTA=0).

The pair kernel accumulates every output element over K in the same order as the one-CTA kernel (same 16-wide MMA
steps, same sequence), so the two must agree BIT FOR BIT; the fp32 reference bounds both."""
import pytest
import torch

pytestmark = [pytest.mark.gpu]
BF16 = torch.bfloat16


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import lia_b200  # noqa: F401
    from lia_b200 import ops as o
    return o


def rnd(*shape, std=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * std).to(BF16).cuda()


# M >= 512 and N % 256 == 0 or N >= 2048 select the pair kernel; ragged M/N exercise the zero-filled halves
SHAPES = [(512, 256, 64), (512, 512, 256), (1024, 768, 512), (8192, 7168, 7168), (777, 2304, 520), (2048, 7168, 28672),
          (640, 50272, 256)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("epilogue", [0, 1, 2])
def test_pair_gemm_matches_one_cta_kernel(ops, monkeypatch, M, N, K, epilogue):
    a = rnd(M, K, seed=M + N)
    w = rnd(N, K, std=K ** -0.5, seed=K + 1)
    bias = rnd(N, std=0.5, seed=5)
    res = rnd(M, N, seed=6) if epilogue == 2 else None
    monkeypatch.setenv("LIA_GEMM_2CTA", "0")
    y1 = ops.gemm(a, w, bias, epilogue=epilogue, residual=res)
    monkeypatch.setenv("LIA_GEMM_2CTA", "1")
    for rep in range(2):
        y2 = ops.gemm(a, w, bias, epilogue=epilogue, residual=res)
        torch.cuda.synchronize()
        ndiff = int((y1.view(torch.int16) != y2.view(torch.int16)).sum())
        assert ndiff == 0, f"{M}x{N}x{K} epi {epilogue} rep {rep}: {ndiff} of {y1.numel()} elements differ from the one-CTA kernel"
    ref = a.float() @ w.float().t()
    err = (y2.float() - (ref.to(BF16).float() + bias.float())).abs().max().item() if epilogue == 0 else 0.0
    assert err <= 0.02 * ref.abs().max().item() + 1e-2


def test_pair_gemm_qkv_scatter_and_model(ops, monkeypatch):
    """The fused QKV epilogue (KV-cache scatter) and a whole prefill through generate(): same tokens, same cache."""
    import lia_b200
    cfg = lia_b200.modeling_opt.get_config("opt-1.3b")
    cfg.num_hidden_layers = 2
    ids = torch.randint(3, cfg.vocab_size, (8, 128), generator=torch.Generator().manual_seed(3))   # M = 1024 rows
    outs, caches = [], []
    for flag in ("0", "1"):
        monkeypatch.setenv("LIA_GEMM_2CTA", flag)
        m = lia_b200.OPTForCausalLM(cfg, "cuda").init_weights(seed=2, bias_std=0.02, ln_std=0.05)
        outs.append(m.generate(ids, max_new_tokens=4, min_new_tokens=4))
        st = next(iter(m._states.values()))
        caches.append([k.clone() for k in st.kc])
        del m
    assert torch.equal(outs[0], outs[1])
    for k0, k1 in zip(*caches):
        assert torch.equal(k0, k1)
