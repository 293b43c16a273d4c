"""Tensor-parallel data path on real hardware.

* ``test_fused_allreduce_two_ranks_one_gpu``: the fused projection + all-reduce + residual kernel
  (lia_gemm_allreduce_bf16) with ``world`` simulated ranks sharing ONE GPU -- each rank's launch is
  capped to a slice of the SMs (LIA_GEMM_MAX_CTAS) and runs on its own stream, so the launches are
  co-resident and really exchange tiles/flags through their arenas.  Checks the one-shot (M <= 128)
  and two-shot (M > 128) protocols, epoch/parity reuse over repeated calls and CUDA-graph replay,
  bit-identical results on every rank, and the reference's rounding points (D:60-77, D:247).
* ``test_tp_multi_gpu``: needs >= 2 GPUs (skipped on a 1-GPU box): torchrun tests/tp_worker.py.
"""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref(As, Ws, bs, res):
    """bf16(residual + bf16(sum_r bf16(bf16(A_r W_r^T) + b_r))), fp32 sums in rank order."""
    tot = None
    for a, w, b in zip(As, Ws, bs):
        part = (a.float() @ w.float().t()).to(BF16)
        part = (part.float() + b.float()).to(BF16).float()
        tot = part if tot is None else tot + part
    return (res.float() + tot.to(BF16).float()).to(BF16)


@pytest.mark.parametrize("world,M,N,K,twoshot", [
    (2, 8, 256, 512, None), (2, 64, 1024, 1536, None), (4, 64, 512, 256, "0"),
    # decode two-shot (default for world >= 4): tile ta is reduced by rank ta % world, finals pushed to all
    (4, 64, 512, 256, None), (2, 64, 1024, 1536, "1"), (4, 24, 1280, 192, None), (8, 64, 2048, 128, None),
    (8, 128, 1024, 512, None), (3, 40, 640, 320, "1"),
    (2, 384, 512, 256, None), (2, 1000, 768, 320, None), (4, 512, 1024, 128, None)])
def test_fused_allreduce_two_ranks_one_gpu(world, M, N, K, twoshot, monkeypatch):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    if twoshot is not None:
        monkeypatch.setenv("LIA_TP_DECODE_TWOSHOT", twoshot)
    import lia_b200  # noqa: F401
    from lia_b200 import _lib, graphs, ops, tp
    lib = _lib.load()
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    monkeypatch.setenv("LIA_GEMM_MAX_CTAS", str(sms // world))
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(world * 1000 + M)
    rnd = lambda *s: (torch.randn(*s, generator=g, device=dev) * 0.5).to(BF16)  # noqa: E731
    As = [rnd(M, K) for _ in range(world)]
    Ws = [rnd(N, K) for _ in range(world)]
    bs = [rnd(N) for _ in range(world)]
    res = rnd(M, N)
    recv = lib.lia_tp_recv_bytes(M, N, K, world)
    arenas = [tp.PeerArena(r, world, dev, recv, [("out", M * N * 2)], exchange=lambda mine: [0] * world)
              for r in range(world)]
    for a in arenas:
        for r in range(world):
            a.peers[r] = arenas[r].local
    outs = [a.tensor("out", (M, N)) for a in arenas]
    wss = [ops.GemmWorkspace(ops.GemmWorkspace.bytes_for([(M, N, K)]), dev) for _ in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    want = _ref(As, Ws, bs, res)

    def launch_all():
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                ops.gemm_allreduce(As[r], Ws[r], bs[r], res, outs[r], arenas[r].args(outs[r] if M > 128 else None),
                                   workspace=wss[r])

    torch.cuda.synchronize()
    for rep in range(3):                       # epochs 1..3: both receive parities, flag reuse
        for o in outs:
            o.zero_()
        torch.cuda.synchronize()
        launch_all()
        torch.cuda.synchronize()
        for a in arenas:
            a.check()
        for r in range(1, world):
            assert torch.equal(outs[0], outs[r]), f"rank {r} differs from rank 0 (rep {rep})"
        err = (outs[0].float() - want.float()).abs().max().item()
        scale = want.float().abs().max().item()
        assert err <= 2 ** -7 * scale, (rep, err, scale)      # a couple of bf16 ulps (K-loop summation order only)
    first = outs[0].clone()
    # CUDA-graph replay: kernel arguments are frozen, epochs advance in device memory
    captured = []
    for r in range(world):
        gr = torch.cuda.CUDAGraph()
        with graphs.capture(gr, stream=streams[r]):       # GC-quiesced capture that surfaces errors raised inside it
            ops.gemm_allreduce(As[r], Ws[r], bs[r], res, outs[r], arenas[r].args(outs[r] if M > 128 else None),
                               workspace=wss[r])
        captured.append(gr)
    for rep in range(2):
        for o in outs:
            o.zero_()
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                captured[r].replay()
        torch.cuda.synchronize()
        for a in arenas:
            a.check()
        for r in range(world):
            assert torch.equal(outs[r], first), f"graph replay {rep}: rank {r} differs"
    for a in arenas:
        a._opened = []
        a.close(sync=False)


def test_finalizers_do_not_invalidate_capture():
    """GPUTEST_r01's failure mode: Python's cyclic GC fires in the middle of somebody's graph capture and a dead
    resource holder's finalizer calls cudaFreeHost / cudaFree / cudaStreamDestroy -> cudaErrorStreamCaptureInvalidated
    at capture_end.  Finalizers must park their frees while a capture is open (lia_b200/graphs.py), also for captures
    opened with plain torch.cuda.graph."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import gc
    import lia_b200  # noqa: F401
    from lia_b200 import graphs, ops, tp
    from lia_b200.kv_spill import KVSpill
    from lia_b200.streamer import HostArena, LayerStreamer
    from lia_b200.weights import LayerLayout
    dev = torch.device("cuda:0")

    def make_garbage():
        lay = LayerLayout(64, 256, 1, heads=1)
        host = HostArena(lay.numel * 2)
        slabs = [host.tensor[i * lay.numel:(i + 1) * lay.numel] for i in range(2)]
        st = LayerStreamer(lay, slabs, dev)
        st.begin()
        sp = KVSpill(2, 8, 2, 1, 64, dev)
        ar = tp.PeerArena(0, 2, dev, 4096, [("out", 4096)], exchange=lambda mine: [0, 0])
        cyc = [host, st, sp, ar]
        cyc.append(cyc)                                   # only the cyclic collector can free these

    x = torch.randn(64, 256, device=dev).to(BF16)
    w = torch.randn(256, 256, device=dev).to(BF16)
    b = torch.zeros(256, device=dev, dtype=BF16)
    out = torch.empty(64, 256, device=dev, dtype=BF16)
    ws = ops.GemmWorkspace(ops.GemmWorkspace.bytes_for([(64, 256, 256)]), dev)
    want = ops.gemm(x, w, b, workspace=ws).clone()
    torch.cuda.synchronize()
    gc.collect()
    for how in ("torch", "ours"):
        make_garbage()
        g = torch.cuda.CUDAGraph()
        if how == "torch":                                # a foreign capture: only the finalizers' own check protects it
            with torch.cuda.graph(g):
                gc.collect()                              # fires the finalizers inside the capture
                ops.gemm(x, w, b, out=out, workspace=ws)
            assert graphs._deferred, "finalizers ran inside the capture instead of being parked"
            graphs.drain()
        else:
            with graphs.capture(g):                       # collects before opening, keeps the collector off inside
                ops.gemm(x, w, b, out=out, workspace=ws)
        assert not graphs._deferred
        out.zero_()
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, want), how
    # an exception raised inside the block is what the caller sees, not the capture_end error that follows from it
    g = torch.cuda.CUDAGraph()
    with pytest.raises(ZeroDivisionError):
        with graphs.capture(g):
            ops.gemm(x, w, b, out=out, workspace=ws)
            1 / 0
    torch.cuda.synchronize()


@pytest.mark.parametrize("nproc", [2])
def test_tp_multi_gpu(nproc):
    if not torch.cuda.is_available() or torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), os.path.join(ROOT, "tests", "tp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    assert "TP_WORKER_OK" in r.stdout, r.stdout[-4000:]
