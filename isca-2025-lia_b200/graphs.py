"""CUDA-graph capture that nothing outside the captured region can invalidate.

A capture started in torch's default ``capture_error_mode="global"`` is invalidated by ANY "potentially unsafe" CUDA call
made while it is open -- cudaFree, cudaFreeHost, cudaStreamDestroy, cudaDeviceSynchronize -- from any code at all.
Python's cyclic garbage collector can run between any two bytecodes, so a dead model still waiting for collection (its
``HostArena`` / ``LayerStreamer`` / ``KVSpill`` / ``PeerArena`` finalizers free pinned memory, streams and peer mappings)
kills the capture that happens to be open when the collector fires; torch >= 2.9 no longer collects before capturing
(torch/cuda/graphs.py ``__enter__``).  The error then surfaces only at ``capture_end`` as
cudaErrorStreamCaptureInvalidated and masks whatever was raised inside the ``with`` block.

Three independent defences:
  * ``capture(graph)`` below: collect garbage BEFORE the capture opens, keep the collector off while it is open, drain
    deferred frees after it closes, and re-raise an exception from inside the block instead of the capture_end error;
  * finalizers (``__del__``) of every resource holder go through ``finalize()``: while a capture is open they park the
    resource on a list instead of freeing it; ``drain()`` frees them at the next safe point;
  * the C side frees under cudaThreadExchangeStreamCaptureMode(relaxed) (csrc/host.cu), the mechanism the CUDA runtime
    provides for allocators that may be entered during a capture.

The reference has no counterpart: it never captures graphs (every op is an eager PyTorch call).
"""
import contextlib
import gc

import torch

_depth = 0          # captures opened through capture() that are still open
_deferred = []      # zero-argument callables that free a resource, parked while a capture was open


def capturing():
    """True while a capture opened by ``capture()`` is open, or while the current stream is capturing (a capture
    opened by someone else with ``torch.cuda.graph``)."""
    if _depth > 0:
        return True
    try:
        return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
    except Exception:
        return False


def finalize(free):
    """Run ``free()`` now, or park it while a capture is open.  For ``__del__`` methods: never raises."""
    if capturing():
        _deferred.append(free)
        return
    try:
        free()
    except Exception:
        pass


def drain():
    """Free everything parked by ``finalize()``; a no-op while a capture is open."""
    if capturing():
        return
    while _deferred:
        free = _deferred.pop()
        try:
            free()
        except Exception:
            pass


@contextlib.contextmanager
def capture(graph, stream=None, pool=None):
    """``with capture(g): ...`` == ``with torch.cuda.graph(g): ...`` made immune to finalizers (module docstring)."""
    global _depth
    drain()
    gc.collect()                      # dead cycles free their CUDA resources now, not in the middle of the capture
    drain()
    was_enabled = gc.isenabled()
    gc.disable()
    _depth += 1
    inner = None
    kw = {}
    if stream is not None:
        kw["stream"] = stream
    if pool is not None:
        kw["pool"] = pool
    try:
        try:
            with torch.cuda.graph(graph, **kw):
                try:
                    yield graph
                except BaseException as e:          # noqa: BLE001 -- re-raised below, after capture_end has run
                    inner = e
        except BaseException:
            if inner is None:
                raise
            # capture_end failed BECAUSE the block failed: report the cause, not the symptom
        if inner is not None:
            raise inner
    finally:
        _depth -= 1
        if was_enabled:
            gc.enable()
        drain()
