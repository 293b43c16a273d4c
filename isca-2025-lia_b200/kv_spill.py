"""KV-cache host spill: per-layer K/V that do not fit in HBM live in pinned host memory.

The reference keeps the WHOLE cache on the host and moves it per minibatch (lia/modeling_opt.py:326-349
``load_kv_cache`` / ``store_cache`` / ``store_cache_decoding``, driven at :1379-1491), allocating fresh pinned
tensors on every call.  Here the cache is HBM-resident wherever it fits (attention then reads it in place);
only the layers beyond the HBM budget -- BASELINE.json config 3, OPT-30B at batch 512: 203 GB of K/V against
180 GB of HBM -- are spilled, and they are the LAST layers of the stack so their first transfer hides under the
resident layers' compute.  Because the cache is time-major ``[Tmax, B, H, d]`` (attentions.py:471-472), the rows
``[0, T)`` of a layer are ONE contiguous block: a spilled layer costs one H2D copy per K and V before its
attention and one D2H copy of just the rows it appended (all ``S`` rows in prefill, one row per decode step),
both on a private stream, double-buffered over two device slots against the compute stream with events --
no device-wide synchronisation, no per-call allocation.
"""
import torch

from . import graphs
from .streamer import HostArena

BF16 = torch.bfloat16


class KVSpill:
    """Two device KV slots + one pinned arena holding K and V of ``n`` spilled layers."""

    def __init__(self, n, Tmax, B, Hl, d, device):
        self.n, self.Tmax, self.B = n, Tmax, B
        self.device = torch.device(device)
        self.shape = (Tmax, B, Hl, d)
        self.row = B * Hl * d                       # elements per cached position
        per = Tmax * self.row
        self.arena = self._new_host(2 * n * per)
        self.arena.tensor.zero_()
        flat = self.arena.tensor
        self.host_k = [flat[(2 * j) * per:(2 * j + 1) * per].view(self.shape) for j in range(n)]
        self.host_v = [flat[(2 * j + 1) * per:(2 * j + 2) * per].view(self.shape) for j in range(n)]
        self.n_slots = min(2, n)
        self.slot_k = [torch.zeros(self.shape, dtype=BF16, device=self.device) for _ in range(self.n_slots)]
        self.slot_v = [torch.zeros(self.shape, dtype=BF16, device=self.device) for _ in range(self.n_slots)]
        self.stream = self._new_stream()
        self.ready = [self._new_event() for _ in range(self.n_slots)]        # H2D into the slot has landed
        self.computed = [self._new_event() for _ in range(self.n_slots)]     # compute no longer touches the slot
        self.loaded = [None] * self.n_slots          # (spilled-layer index, rows present) per slot
        # --no-overlap (lia/modeling_opt.py:1173): every spilled layer passes through slot 0 and nothing is fetched
        # ahead, so a layer's H2D copy starts only when the previous spilled layer's kernels have released the slot --
        # transfer and compute serialise, which is what the reference's ablation flag measures
        self.overlap = True
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    @property
    def slot_bytes(self):
        return 2 * self.Tmax * self.row * 2

    # ---- device plumbing (the CPU test of the transfer schedule substitutes simulated streams)
    def _new_host(self, numel):
        return HostArena(numel)

    def _new_stream(self):
        return torch.cuda.Stream(self.device)

    def _new_event(self):
        return torch.cuda.Event()

    def _current_stream(self):
        return torch.cuda.current_stream(self.device)

    def _copy_async(self, dst, src):
        with torch.cuda.stream(self.stream):
            dst.copy_(src, non_blocking=True)

    def _slot(self, j):
        return j % self.n_slots if self.overlap else 0

    # ---- all copies are issued on self.stream, in order: a layer's store always precedes its next load
    def _prefetch(self, j, rows):
        slot = self._slot(j)
        if self.loaded[slot] == (j, rows):
            return
        # the slot's previous tenant must be done with it (its release recorded `computed`; a never-recorded event
        # is a no-op): with overlap the release path has already made the stream wait, without it the tenant changes
        # on every layer
        self.stream.wait_event(self.computed[slot])
        if rows > 0:
            self._copy_async(self.slot_k[slot][:rows], self.host_k[j][:rows])
            self._copy_async(self.slot_v[slot][:rows], self.host_v[j][:rows])
            self.h2d_bytes += 2 * rows * self.row * 2
        self.ready[slot].record(self.stream)
        self.loaded[slot] = (j, rows)

    def begin(self, pos0):
        """Start of a forward that finds ``pos0`` rows cached: make sure the first slots are in flight."""
        if not self.overlap:
            return
        for j in range(self.n_slots):
            self._prefetch(j, pos0)

    def acquire(self, j, pos0):
        """(K, V) device views of spilled layer j holding rows [0, pos0), valid on the current stream."""
        slot = self._slot(j)
        if not self.overlap and self.loaded[slot] != (j, pos0):
            self.computed[slot].record(self._current_stream())   # the copy starts only after everything enqueued so far
        self._prefetch(j, pos0)
        self._current_stream().wait_event(self.ready[slot])
        return self.slot_k[slot], self.slot_v[slot]

    def release(self, j, pos0, S):
        """Layer j's kernels (which appended rows [pos0, pos0+S)) are enqueued: write those rows back to the
        host copy, then recycle the slot for layer j+2 -- wrapping around into the NEXT forward, which will
        find pos0+S rows."""
        slot = self._slot(j)
        self.computed[slot].record(self._current_stream())
        self.stream.wait_event(self.computed[slot])
        self._copy_async(self.host_k[j][pos0:pos0 + S], self.slot_k[slot][pos0:pos0 + S])
        self._copy_async(self.host_v[j][pos0:pos0 + S], self.slot_v[slot][pos0:pos0 + S])
        self.d2h_bytes += 2 * S * self.row * 2
        self.loaded[slot] = (j, pos0 + S)             # the slot now holds this layer with the new rows
        if self.n <= self.n_slots or not self.overlap:
            return                                    # every spilled layer owns a slot (or no fetching ahead): nothing to recycle
        nxt, rows = j + self.n_slots, pos0
        if nxt >= self.n:                             # first layers of the NEXT forward (odd counts: begin() fetches them)
            nxt, rows = nxt - self.n, pos0 + S
            if rows >= self.Tmax - 1:                 # the cache is full: no next step (greedy_search.py:425)
                return
        if nxt % self.n_slots == slot:
            self._prefetch(nxt, rows)

    def synchronize(self):
        self.stream.synchronize()

    def stats(self):
        return {"layers": self.n, "host_bytes": self.arena.nbytes, "h2d_bytes": self.h2d_bytes, "d2h_bytes": self.d2h_bytes}

    def close(self):
        if self.arena is not None:
            self.synchronize()
            self.host_k = self.host_v = None
            self.arena.close()
            self.arena = None

    def __del__(self):
        # synchronising a stream / freeing pinned memory inside an open graph capture would invalidate it (graphs.py)
        if getattr(self, "arena", None) is not None:
            graphs.finalize(self.close)


def plan_resident_layers(L, per_layer_bytes, free_bytes, reserve_bytes=4 << 30):
    """How many layers' K/V stay in HBM: all of them if they fit next to ``reserve_bytes`` of head-room,
    otherwise as many as fit once the two spill slots are paid for."""
    if L * per_layer_bytes + reserve_bytes <= free_bytes:
        return L
    n = (free_bytes - reserve_bytes - 2 * per_layer_bytes) // per_layer_bytes
    return int(max(0, min(L - 1, n)))
