"""KV-cache host spill (reference: load_kv_cache / store_cache / store_cache_decoding,
lia/modeling_opt.py:326-349): the transfer SCHEDULE checked on CPU against simulated CUDA streams.

The simulator keeps one FIFO per stream and runs them in random interleavings that respect only what CUDA
guarantees (in-order per stream; ``wait_event`` blocks on the most recent ``record`` made before it), so a
missing event shows up as a layer reading another layer's rows or the host copy missing a row."""
import os
import random
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lia_b200  # noqa: E402,F401
from lia_b200.kv_spill import KVSpill, plan_resident_layers  # noqa: E402

BF16 = torch.bfloat16


class SimStream:
    def __init__(self, sim, name):
        self.sim, self.name, self.q = sim, name, []
        sim.streams.append(self)

    def enqueue(self, fn):
        self.q.append(("run", fn))

    def wait_event(self, ev):
        if ev.last is not None:
            self.q.append(("wait", ev.last))

    def synchronize(self):
        self.sim.drain()


class SimEvent:
    def __init__(self, sim):
        self.sim, self.last = sim, None

    def record(self, stream):
        self.sim.serial += 1
        self.last = self.sim.serial
        stream.q.append(("mark", self.last))


class Sim:
    def __init__(self, seed):
        self.rng = random.Random(seed)
        self.streams, self.done, self.serial = [], set(), 0

    def step(self):
        runnable = [s for s in self.streams if s.q and not (s.q[0][0] == "wait" and s.q[0][1] not in self.done)]
        if not runnable:
            assert not any(s.q for s in self.streams), "deadlock: a stream waits on an event that is never recorded"
            return False
        s = self.rng.choice(runnable)
        kind, arg = s.q.pop(0)
        if kind == "run":
            arg()
        elif kind == "mark":
            self.done.add(arg)
        return True

    def drain(self):
        while self.step():
            pass


class _FakeArena:
    def __init__(self, numel):
        self.tensor = torch.empty(numel, dtype=BF16)
        self.nbytes = numel * 2

    def close(self):
        self.tensor = None


class SimKVSpill(KVSpill):
    def __init__(self, sim, *a):
        self.sim = sim
        self.compute = SimStream(sim, "compute")
        super().__init__(*a)

    def _new_host(self, numel):
        return _FakeArena(numel)

    def _new_stream(self):
        return SimStream(self.sim, "copy")

    def _new_event(self):
        return SimEvent(self.sim)

    def _current_stream(self):
        return self.compute

    def _copy_async(self, dst, src):
        self.stream.enqueue(lambda: dst.copy_(src))


def val(layer, row, kv):
    """Distinct bf16-exact value per (layer, cache row, K|V): below 128 the bf16 grid is 0.5 or finer."""
    assert row < 9 and layer < 14
    return float(layer * 9 + row + (0.5 if kv else 0.0))


def run_generation(n, S, new, seed, lazy_steps=3, cls=None, overlap=True):
    """Prefill of S rows then new-1 decode steps over n spilled layers, the way OPTDecoder.run_layers drives
    the spill.  Each layer's "kernel" checks that the slot holds exactly its own history and appends its rows."""
    sim = Sim(seed)
    Tmax, B, Hl, d = S + new, 2, 1, 8
    sp = (cls or SimKVSpill)(sim, n, Tmax, B, Hl, d, "cpu")
    sp.overlap = overlap
    errors = []

    def layer_kernel(j, kc, vc, pos0, rows):
        def fn():
            for r in range(pos0):
                if kc[r, 0, 0, 0].item() != val(j, r, 0) or vc[r, 0, 0, 0].item() != val(j, r, 1):
                    errors.append((j, pos0, r, kc[r, 0, 0, 0].item(), val(j, r, 0)))
                    break
            for r in range(pos0, pos0 + rows):
                kc[r].fill_(val(j, r, 0))
                vc[r].fill_(val(j, r, 1))
        return fn

    def forward(pos0, rows):
        sp.begin(pos0)
        for j in range(n):
            kc, vc = sp.acquire(j, pos0)
            sp.compute.enqueue(layer_kernel(j, kc, vc, pos0, rows))
            sp.release(j, pos0, rows)
            for _ in range(sim.rng.randrange(lazy_steps)):      # the device runs ahead of / behind the host at random
                sim.step()

    for _rep in range(2):                                       # a second generate() re-uses the same state
        forward(0, S)
        for t in range(1, new):
            forward(S + t - 1, 1)
        sp.synchronize()
        assert not errors, errors[:3]
        for j in range(n):
            for r in range(Tmax - 1):
                assert sp.host_k[j][r, 1, 0, 3].item() == val(j, r, 0), (j, r)
                assert sp.host_v[j][r, 0, 0, 7].item() == val(j, r, 1), (j, r)
    return sp


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 12])
def test_spill_schedule_under_random_interleavings(n):
    for seed in range(12):
        sp = run_generation(n, S=5, new=4, seed=seed)
    if n <= 2:              # every spilled layer owns a slot: after the first load nothing is re-read from the host
        assert sp.h2d_bytes == 0
    else:                   # each decode step re-reads each layer's history once (S + t - 1 rows)
        per_row = 2 * sp.row * 2
        want = 2 * n * sum(5 + t - 1 for t in range(1, 4)) * per_row
        slack = 2 * 2 * 8 * per_row                       # a wrap-around prefetch the next generate() does not use
        assert sp.h2d_bytes <= want + slack, (sp.h2d_bytes, want)
        if n % 2 == 0:                                    # (odd counts: one layer keeps its slot and is never re-read)
            assert sp.h2d_bytes >= want, (sp.h2d_bytes, want)
    assert sp.d2h_bytes == 2 * n * (5 + 3) * 2 * sp.row * 2   # exactly the appended rows go back, never the history


class SerialCheckSpill(SimKVSpill):
    """An H2D copy (destination inside a device slot) asserts, when it executes, that the compute stream has nothing it
    could be running at the same time: its queue is empty or its head is a wait that has not been satisfied yet."""

    def __init__(self, sim, *a):
        super().__init__(sim, *a)
        self.slot_storages = {t.untyped_storage().data_ptr() for t in self.slot_k + self.slot_v}

    def _copy_async(self, dst, src):
        h2d = dst.untyped_storage().data_ptr() in self.slot_storages

        def fn():
            q = self.compute.q
            if h2d:
                assert not q or (q[0][0] == "wait" and q[0][1] not in self.sim.done), "an H2D copy ran alongside a runnable kernel"
            dst.copy_(src)
        self.stream.enqueue(fn)


@pytest.mark.parametrize("n", [1, 2, 3, 5])
def test_no_overlap_schedule_is_correct_and_serial(n):
    """--no-overlap (lia/modeling_opt.py:1173): every spilled layer goes through slot 0, nothing is fetched ahead, and a
    layer's H2D copy is ordered after every kernel enqueued before its acquire."""
    for seed in range(12):
        sp = run_generation(n, S=5, new=4, seed=seed, cls=SerialCheckSpill, overlap=False)
        assert sp.loaded[0] is not None and all(x is None for x in sp.loaded[1:])      # only slot 0 is ever used
    per_row = 2 * sp.row * 2
    if n == 1:
        assert sp.h2d_bytes == 0                          # the one spilled layer keeps the slot
    else:                                                 # every layer re-reads its history on every decode step, both passes
        assert sp.h2d_bytes == 2 * n * sum(5 + t - 1 for t in range(1, 4)) * per_row
    assert sp.d2h_bytes == 2 * n * (5 + 3) * 2 * sp.row * 2
    # the check itself has teeth: with overlap the same harness sees copies running ahead of pending kernels
    with pytest.raises(AssertionError, match="H2D copy ran alongside"):
        for seed in range(40):
            run_generation(4, S=5, new=4, seed=seed, cls=SerialCheckSpill, overlap=True)


def test_simulator_catches_a_missing_event():
    """The harness itself: drop the wait on `ready` and the hazard is detected for some interleaving."""
    class Broken(SimKVSpill):
        def acquire(self, j, pos0):
            slot = j % self.n_slots
            self._prefetch(j, pos0)
            return self.slot_k[slot], self.slot_v[slot]

    with pytest.raises(AssertionError):
        for seed in range(40):
            run_generation(4, S=5, new=4, seed=seed, cls=Broken)


def test_plan_resident_layers():
    GB = 1 << 30
    per = int(4.23 * GB)                                           # OPT-30B, B=512, T=288: K+V of one layer
    assert plan_resident_layers(48, int(0.53 * GB), 100 * GB) == 48    # config 2 (B=64): everything fits
    n = plan_resident_layers(48, per, 150 * GB)                        # config 3: 203 GB of K/V > HBM
    assert 0 < n < 48 and (n + 2) * per + 4 * GB <= 150 * GB < (n + 3) * per + 4 * GB
    assert plan_resident_layers(48, per, 2 * GB) == 0
    assert plan_resident_layers(1, per, 2 * GB) == 0
