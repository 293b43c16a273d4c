// Shared host/device helpers for libliab200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "lia_b200.h"

// ---------------------------------------------------------------- host-side error plumbing
void lia_set_error(const char* fmt, ...);   // c_abi.cu

#define LIA_CHECK_ARG(cond, ...)          \
  do {                                    \
    if (!(cond)) {                        \
      lia_set_error(__VA_ARGS__);         \
      return LIA_ERR_INVALID;             \
    }                                     \
  } while (0)

#define LIA_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      lia_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return LIA_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

#define LIA_LAUNCH_CHECK()                                                                     \
  do {                                                                                         \
    cudaError_t e__ = cudaGetLastError();                                                      \
    if (e__ != cudaSuccess) {                                                                  \
      lia_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return LIA_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

int lia_sm_count();   // cached, c_abi.cu
bool lia_pdl_enabled();   // LIA_PDL=0 disables programmatic dependent launch, c_abi.cu

#ifdef __CUDACC__
#include <utility>
// Every kernel is launched with programmatic stream serialization allowed (PDL): the next kernel's
// CTAs may become resident and run their prologue (and, in the GEMM, prefetch weight tiles) while
// the previous kernel drains; each kernel calls pdl_wait() before it touches anything a
// predecessor produced.  Works unchanged inside CUDA-graph capture.
template <typename... KArgs, typename... Args>
inline cudaError_t lia_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = lia_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#endif

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__

typedef __nv_bfloat16 bf16;

// round-to-nearest-even to bf16 and back: one "rounding point" of the reference's eager ops
__device__ __forceinline__ float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void unpack_bf16x2(uint32_t u, float& lo, float& hi) {
  lo = __uint_as_float(u << 16);
  hi = __uint_as_float(u & 0xffff0000u);
}
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  unpack_bf16x2(u.x, f[0], f[1]);
  unpack_bf16x2(u.y, f[2], f[3]);
  unpack_bf16x2(u.z, f[4], f[5]);
  unpack_bf16x2(u.w, f[6], f[7]);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]);
  u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

// streaming 128-bit loads that do not pollute L1 (data read once: KV cache, activations)
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// 128-bit load of ACTIVATIONS: cached in L2 only, so it always sees what other SMs have written -- also data produced
// earlier in the SAME launch (the decode program kernel chains operations inside one grid; a non-coherent .nc load could
// return a stale L1 line there).  Costs the same as ldg_stream: neither allocates in L1.
__device__ __forceinline__ uint4 ldg_act(const void* p) {
  uint4 r;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// programmatic dependent launch (no-ops when the kernel was launched without the attribute)
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#endif  // __CUDACC__
