"""Pinned-host layer streaming: the B200 replacement for LIA's AMX-CPU compute + CXL tiering.

Layers that are not HBM-resident (``--gpu-percentage`` < 100, or a model larger than 180 GB)
live as contiguous slabs in a pinned host arena and are copied, one cudaMemcpyAsync per layer,
into two device slots on a private copy stream, double-buffered against the compute stream
(reference: load_layer/layer_copy + 5 streams + cuda.synchronize per minibatch,
lia/modeling_opt.py:270-318, 1208-1212, 1288-1339; pin_memory / CXL alloc :167-227).
"""
import ctypes

import torch

from . import _lib, graphs
from ._lib import check

BF16 = torch.bfloat16


class HostArena:
    """One pinned allocation from lia_host_arena_alloc, exposed as a flat bf16 tensor.
    Mirrors lia/cxl/numa_alloc.py:28-50 (raw pointer wrapped into a torch tensor)."""

    def __init__(self, numel):
        self.nbytes = int(numel) * 2
        self.ptr = _lib.load().lia_host_arena_alloc(self.nbytes)
        if not self.ptr:
            raise MemoryError(f"pinned host arena of {self.nbytes} bytes: {_lib.last_error()}")   # as M:175
        buf = (ctypes.c_uint8 * self.nbytes).from_address(self.ptr)
        self.tensor = torch.frombuffer(buf, dtype=BF16, count=int(numel))

    def close(self):
        if self.ptr:
            self.tensor = None
            ptr, self.ptr = self.ptr, None
            check(_lib.load().lia_host_arena_free(ptr, self.nbytes), "lia_host_arena_free")

    def __del__(self):
        # cudaFreeHost inside somebody's open graph capture would invalidate it: park the free (graphs.py)
        ptr, nbytes = getattr(self, "ptr", None), getattr(self, "nbytes", 0)
        if ptr:
            self.ptr, self.tensor = None, None
            graphs.finalize(lambda: _lib.load().lia_host_arena_free(ptr, nbytes))


class LayerStreamer:
    """Two device slots + the copy stream; ``layers`` are host slabs (flat bf16 tensors)."""

    def __init__(self, layout, host_slabs, device):
        self.layout = layout
        self.host = host_slabs
        self.n = len(host_slabs)
        self.n_slots = min(2, self.n)
        self.slots = [torch.empty(layout.numel, dtype=BF16, device=device) for _ in range(self.n_slots)]
        self.views = [layout.views(s) for s in self.slots]
        arr = (ctypes.c_void_p * self.n_slots)(*[s.data_ptr() for s in self.slots])
        self.handle = _lib.load().lia_streamer_create(arr, self.n_slots, layout.nbytes)
        if not self.handle:
            raise _lib.LiaError(f"lia_streamer_create: {_lib.last_error()}")
        self.loaded = [-1] * self.n_slots     # which streamed-layer index each slot holds / is receiving
        # --no-overlap (lia/modeling_opt.py:1173, load_layer(..., overlap=False)): every streamed layer passes through
        # slot 0 and nothing is fetched ahead; lia_streamer_prefetch waits for the slot's release, i.e. for the previous
        # streamed layer's kernels, so copy and compute serialise -- the reference's ablation of its weight prefetch
        self.overlap = True

    def _slot(self, j):
        return j % self.n_slots if self.overlap else 0

    def _prefetch(self, j):
        slot = self._slot(j)
        if self.loaded[slot] == j:
            return
        check(_lib.load().lia_streamer_prefetch(self.handle, slot, self.host[j].data_ptr(), self.layout.nbytes),
              "lia_streamer_prefetch")
        self.loaded[slot] = j

    def begin(self):
        if not self.overlap:
            return
        for j in range(self.n_slots):
            self._prefetch(j)

    def acquire(self, j):
        """Views of streamed layer j, valid on the current stream after this call."""
        slot = self._slot(j)
        if not self.overlap and self.loaded[slot] != j:
            # the copy starts only after everything enqueued so far on the compute stream (prefetch waits for the release)
            check(_lib.load().lia_streamer_release(self.handle, slot, torch.cuda.current_stream().cuda_stream),
                  "lia_streamer_release")
        self._prefetch(j)
        check(_lib.load().lia_streamer_wait(self.handle, slot, torch.cuda.current_stream().cuda_stream),
              "lia_streamer_wait")
        return self.views[slot]

    def release(self, j):
        """Layer j's compute has been enqueued: recycle its slot for layer j+2 (wrapping around so
        the next forward finds its first layers already in flight)."""
        slot = self._slot(j)
        check(_lib.load().lia_streamer_release(self.handle, slot, torch.cuda.current_stream().cuda_stream),
              "lia_streamer_release")
        if self.n > self.n_slots and self.overlap:
            nxt = j + self.n_slots
            if nxt >= self.n:
                nxt -= self.n                      # first layers of the NEXT forward
            if nxt % self.n_slots == slot:         # (odd layer counts: begin() fetches them instead)
                self._prefetch(nxt)

    def stats(self):
        b, ms = ctypes.c_double(0), ctypes.c_double(0)
        check(_lib.load().lia_streamer_stats(self.handle, ctypes.byref(b), ctypes.byref(ms)), "lia_streamer_stats")
        return {"bytes": b.value, "copy_ms": ms.value,
                "gbps": (b.value / 1e9) / (ms.value / 1e3) if ms.value > 0 else 0.0}

    def close(self):
        if self.handle:
            handle, self.handle = self.handle, None
            _lib.load().lia_streamer_destroy(handle)

    def __del__(self):
        handle = getattr(self, "handle", None)
        if handle:
            self.handle = None
            slots = getattr(self, "slots", None)

            def free(slots=slots):                        # the device slots must outlive the copy stream
                _lib.load().lia_streamer_destroy(handle)
                del slots
            graphs.finalize(free)
