// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/vol_latency scripts/vol_latency.cu
// What does one round of the decode exchange's polling cost?  148 CTAs x 128 threads, each thread issues G 16-byte loads
// (all in flight), waits for them, repeats R dependent rounds; and S 16-byte stores.  Flavours: ld.volatile (what the LL
// protocol uses), ld.relaxed.gpu, ld.global.cg (weak, L2); st.volatile vs st.global (weak).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint4 ld_vol(const uint4* p) { uint4 v; asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ uint4 ld_rlx(const uint4* p) { uint4 v; asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ uint4 ld_cg(const uint4* p) { uint4 v; asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_vol(uint4* p, uint4 v) { asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__device__ __forceinline__ void st_weak(uint4* p, uint4 v) { asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }

template <int MODE, int G>
__global__ void loads(const uint4* buf, size_t stride, int rounds, long long* out, uint32_t* sink) {
  const uint4* p = buf + (size_t)blockIdx.x * 4096 + threadIdx.x * 2;
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    uint4 v[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const uint4* q = p + (size_t)g * stride + (size_t)(r & 3) * 256 + (acc & 1);
      v[g] = MODE == 0 ? ld_vol(q) : MODE == 1 ? ld_rlx(q) : ld_cg(q);
    }
#pragma unroll
    for (int g = 0; g < G; ++g) acc += v[g].x + v[g].w;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) *sink = acc;
}
template <int MODE>
__global__ void stores(uint4* buf, size_t stride, int n, long long* out) {
  uint4* p = buf + (size_t)blockIdx.x * 4096 + threadIdx.x * 2;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    uint4 v = make_uint4(i, 1, i, 1);
    if (MODE == 0) st_vol(p + (size_t)(i & 7) * stride + (i >> 3), v); else st_weak(p + (size_t)(i & 7) * stride + (i >> 3), v);
  }
  long long t1 = clock64();
  __syncthreads();
  long long t2 = clock64();
  if (threadIdx.x == 0) { out[blockIdx.x] = t1 - t0; out[blockIdx.x + 256] = t2 - t0; }
}
int main() {
  const size_t stride = 64 * 7168 * 2 / 16 * 2;   // one source slot of the receive area, in uint4
  const size_t n = stride * 10 + 149 * 4096 + 4096;
  uint4* buf; cudaMalloc(&buf, n * 16); cudaMemset(buf, 1, n * 16);
  cudaIpcMemHandle_t h; cudaIpcGetMemHandle(&h, buf);
  long long *out, hout[512]; cudaMalloc(&out, 512 * 8); uint32_t* sink; cudaMalloc(&sink, 4);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  auto report = [&](const char* what, int per) {
    cudaDeviceSynchronize(); cudaMemcpy(hout, out, 512 * 8, cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < 148; ++i) s += hout[i];
    printf("%-44s %8.0f cycles per %s\n", what, s / 148 / per, per == 1 ? "call" : "round");
  };
  const int R = 16;
  for (int rep = 0; rep < 2; ++rep) {
    loads<0, 8><<<148, 128>>>(buf, stride, R, out, sink); report("ld.volatile   x8 in flight", R);
    loads<1, 8><<<148, 128>>>(buf, stride, R, out, sink); report("ld.relaxed.gpu x8 in flight", R);
    loads<2, 8><<<148, 128>>>(buf, stride, R, out, sink); report("ld.global.cg  x8 in flight", R);
    loads<0, 2><<<148, 128>>>(buf, stride, R, out, sink); report("ld.volatile   x2 in flight", R);
    loads<2, 2><<<148, 128>>>(buf, stride, R, out, sink); report("ld.global.cg  x2 in flight", R);
    loads<0, 16><<<148, 128>>>(buf, stride / 2, R, out, sink); report("ld.volatile   x16 in flight", R);
    stores<0><<<148, 128>>>(buf, stride, 16, out); report("st.volatile x16 (issue)", 1);
    stores<1><<<148, 128>>>(buf, stride, 16, out); report("st.global   x16 (issue)", 1);
  }
  printf("SM clock attr %d kHz; err=%s\n", clk, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
