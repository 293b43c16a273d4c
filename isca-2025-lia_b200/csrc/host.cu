// Host-side runtime pieces of libliab200: the pinned host arena and the layer streamer.
//
//   lia_host_arena_*  replaces lia/cxl/numa_alloc.c (numa_alloc_node/numa_free_node: raw pointer out,
//                     caller frees with the size) and pin_memory() (lia/modeling_opt.py:167-227).
//   lia_streamer_*    replaces load_layer/layer_copy (lia/modeling_opt.py:270-318: 16 copy_() calls per
//                     layer) and the per-forward stream/buffer churn (:1195-1212, :1288-1316) with ONE
//                     cudaMemcpyAsync per layer slab on a private copy stream, double-buffered against
//                     the compute stream with events -- no device-wide synchronisation.
#include <string.h>

#include <vector>

#include "common.cuh"

// Allocation and release may be entered while SOME stream of the process is capturing a CUDA graph (a Python finalizer
// fired by the garbage collector in the middle of a capture; a second model built while the first one's decode graphs
// are being captured).  Under the default global capture mode such "potentially unsafe" calls (cudaFree, cudaFreeHost,
// cudaMalloc, stream/event destruction) would INVALIDATE that capture; switching this thread to relaxed mode for the
// duration of the call is the runtime's own mechanism for allocators (cudaThreadExchangeStreamCaptureMode).
struct RelaxedCaptureMode {
  cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
  RelaxedCaptureMode() { cudaThreadExchangeStreamCaptureMode(&mode); }
  ~RelaxedCaptureMode() { cudaThreadExchangeStreamCaptureMode(&mode); }
};

extern "C" void* lia_host_arena_alloc(size_t bytes) {
  RelaxedCaptureMode relaxed;
  void* p = nullptr;
  if (bytes == 0) {
    lia_set_error("lia_host_arena_alloc: zero bytes");
    return nullptr;
  }
  cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocPortable);
  if (e != cudaSuccess) {
    lia_set_error("lia_host_arena_alloc: cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    (void)cudaGetLastError();
    return nullptr;
  }
  return p;
}

extern "C" int lia_host_arena_free(void* ptr, size_t /*bytes*/) {
  if (ptr == nullptr) return LIA_OK;
  RelaxedCaptureMode relaxed;
  LIA_CUDA(cudaFreeHost(ptr));
  return LIA_OK;
}

struct LiaStreamer {
  std::vector<void*> slabs;
  size_t slab_bytes = 0;
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> ready;     // copy into slot finished
  std::vector<cudaEvent_t> released;  // compute no longer reads slot
  std::vector<bool> has_release;
  // copy timing: a fixed ring of event pairs, re-used and folded into copy_ms when recycled -- nothing is created on
  // the hot path and nothing accumulates when the caller never asks for statistics
  static constexpr int TIMING_RING = 64;
  cudaEvent_t t0[TIMING_RING] = {};
  cudaEvent_t t1[TIMING_RING] = {};
  bool pending[TIMING_RING] = {};
  int ring_pos = 0;
  double bytes = 0.0;
  double copy_ms = 0.0;
  void fold(int i) {
    if (!pending[i]) return;
    float ms = 0.f;
    if (cudaEventSynchronize(t1[i]) == cudaSuccess && cudaEventElapsedTime(&ms, t0[i], t1[i]) == cudaSuccess) copy_ms += ms;
    pending[i] = false;
  }
};

extern "C" LiaStreamer* lia_streamer_create(void* const* device_slabs, int n_slots, size_t slab_bytes) {
  if (device_slabs == nullptr || n_slots <= 0 || slab_bytes == 0) {
    lia_set_error("lia_streamer_create: bad arguments");
    return nullptr;
  }
  LiaStreamer* s = new (std::nothrow) LiaStreamer();
  if (!s) {
    lia_set_error("lia_streamer_create: out of host memory");
    return nullptr;
  }
  s->slab_bytes = slab_bytes;
  if (cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
    lia_set_error("lia_streamer_create: cudaStreamCreate failed");
    delete s;
    return nullptr;
  }
  for (int i = 0; i < n_slots; ++i) {
    s->slabs.push_back(device_slabs[i]);
    cudaEvent_t a = nullptr, b = nullptr;
    if (cudaEventCreateWithFlags(&a, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&b, cudaEventDisableTiming) != cudaSuccess) {
      lia_set_error("lia_streamer_create: cudaEventCreate failed");
      lia_streamer_destroy(s);
      return nullptr;
    }
    s->ready.push_back(a);
    s->released.push_back(b);
    s->has_release.push_back(false);
  }
  for (int i = 0; i < LiaStreamer::TIMING_RING; ++i) {
    if (cudaEventCreate(&s->t0[i]) != cudaSuccess || cudaEventCreate(&s->t1[i]) != cudaSuccess) {
      lia_set_error("lia_streamer_create: cudaEventCreate failed");
      lia_streamer_destroy(s);
      return nullptr;
    }
  }
  return s;
}

extern "C" int lia_streamer_prefetch(LiaStreamer* s, int slot, const void* host_src, size_t bytes) {
  LIA_CHECK_ARG(s && slot >= 0 && slot < (int)s->slabs.size(), "lia_streamer_prefetch: bad slot");
  LIA_CHECK_ARG(host_src && bytes > 0 && bytes <= s->slab_bytes, "lia_streamer_prefetch: %zu bytes do not fit the %zu-byte slab", bytes, s->slab_bytes);
  if (s->has_release[slot]) LIA_CUDA(cudaStreamWaitEvent(s->copy_stream, s->released[slot], 0));
  const int ti = s->ring_pos;
  s->ring_pos = (ti + 1) % LiaStreamer::TIMING_RING;
  s->fold(ti);   // the pair's previous copy was 64 prefetches ago: long complete
  LIA_CUDA(cudaEventRecord(s->t0[ti], s->copy_stream));
  LIA_CUDA(cudaMemcpyAsync(s->slabs[slot], host_src, bytes, cudaMemcpyHostToDevice, s->copy_stream));
  LIA_CUDA(cudaEventRecord(s->t1[ti], s->copy_stream));
  LIA_CUDA(cudaEventRecord(s->ready[slot], s->copy_stream));
  s->pending[ti] = true;
  s->bytes += (double)bytes;
  return LIA_OK;
}

extern "C" int lia_streamer_wait(LiaStreamer* s, int slot, lia_stream_t compute_stream) {
  LIA_CHECK_ARG(s && slot >= 0 && slot < (int)s->slabs.size(), "lia_streamer_wait: bad slot");
  LIA_CUDA(cudaStreamWaitEvent(reinterpret_cast<cudaStream_t>(compute_stream), s->ready[slot], 0));
  return LIA_OK;
}

extern "C" int lia_streamer_release(LiaStreamer* s, int slot, lia_stream_t compute_stream) {
  LIA_CHECK_ARG(s && slot >= 0 && slot < (int)s->slabs.size(), "lia_streamer_release: bad slot");
  LIA_CUDA(cudaEventRecord(s->released[slot], reinterpret_cast<cudaStream_t>(compute_stream)));
  s->has_release[slot] = true;
  return LIA_OK;
}

extern "C" int lia_streamer_stats(LiaStreamer* s, double* bytes, double* copy_ms) {
  LIA_CHECK_ARG(s != nullptr, "lia_streamer_stats: null streamer");
  LIA_CUDA(cudaStreamSynchronize(s->copy_stream));
  for (int i = 0; i < LiaStreamer::TIMING_RING; ++i) s->fold(i);
  if (bytes) *bytes = s->bytes;
  if (copy_ms) *copy_ms = s->copy_ms;
  return LIA_OK;
}

extern "C" int lia_streamer_destroy(LiaStreamer* s) {
  if (!s) return LIA_OK;
  RelaxedCaptureMode relaxed;
  if (s->copy_stream) cudaStreamSynchronize(s->copy_stream);
  for (int i = 0; i < LiaStreamer::TIMING_RING; ++i) {
    if (s->t0[i]) cudaEventDestroy(s->t0[i]);
    if (s->t1[i]) cudaEventDestroy(s->t1[i]);
  }
  for (auto e : s->ready) cudaEventDestroy(e);
  for (auto e : s->released) cudaEventDestroy(e);
  if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
  delete s;
  return LIA_OK;
}

// ---------------------------------------------------------------- peer-mappable device arenas (tensor parallelism)
// One cudaMalloc'd arena per rank, exported with CUDA IPC and mapped by every peer process: the fused
// projection + all-reduce kernel (gemm_sm100.cu) stores partial tiles and flags straight into peers'
// arenas over NVLink.  Replaces the reference's oneCCL/MPI/POSIX-shm messenger (csrc/cpu/comm/messager.h:13-62,
// shm_reduction.h) -- there the partial sums travel GPU -> host -> shm/CCL -> host -> GPU (decoder.py:60-68).
static_assert(sizeof(cudaIpcMemHandle_t) == LIA_P2P_HANDLE_BYTES, "LIA_P2P_HANDLE_BYTES must match cudaIpcMemHandle_t");

extern "C" int lia_p2p_alloc(size_t bytes, void** dev_ptr, void* handle_out) {
  LIA_CHECK_ARG(bytes > 0 && dev_ptr != nullptr && handle_out != nullptr, "lia_p2p_alloc: bad arguments");
  RelaxedCaptureMode relaxed;
  void* p = nullptr;
  LIA_CUDA(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    lia_set_error("lia_p2p_alloc(%zu): %s", bytes, cudaGetErrorString(e));
    cudaFree(p);
    (void)cudaGetLastError();
    return LIA_ERR_CUDA;
  }
  memcpy(handle_out, &h, sizeof(h));
  *dev_ptr = p;
  return LIA_OK;
}

extern "C" int lia_p2p_open(const void* handle, void** peer_ptr) {
  LIA_CHECK_ARG(handle != nullptr && peer_ptr != nullptr, "lia_p2p_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  LIA_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *peer_ptr = p;
  return LIA_OK;
}

extern "C" int lia_p2p_close(void* peer_ptr) {
  if (peer_ptr == nullptr) return LIA_OK;
  RelaxedCaptureMode relaxed;
  LIA_CUDA(cudaIpcCloseMemHandle(peer_ptr));
  return LIA_OK;
}

extern "C" int lia_p2p_free(void* dev_ptr) {
  if (dev_ptr == nullptr) return LIA_OK;
  RelaxedCaptureMode relaxed;
  LIA_CUDA(cudaFree(dev_ptr));
  return LIA_OK;
}
