// Isolate what makes back-to-back launches of the GEMM kernel cost ~15 us each beyond the CTA lifetime.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

struct Big { unsigned char b[128]; };

template <bool TMEM, bool BIGPARAM>
__global__ void __launch_bounds__(256, 1) k(const __grid_constant__ Big p1, const __grid_constant__ Big p2, int spin_ns, int* sink) {
  extern __shared__ uint8_t smem[];
  __shared__ uint32_t tptr;
  if (TMEM) {
    if (threadIdx.x / 32 == 2) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tptr)), "r"(128u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
  }
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while (t - t0 < (unsigned long long)spin_ns);
  if (spin_ns < 0) sink[0] = smem[threadIdx.x] + p1.b[0] + p2.b[1];
  if (TMEM) {
    __syncthreads();
    if (threadIdx.x / 32 == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tptr), "r"(128u) : "memory");
  }
}

template <bool TMEM, bool BIGPARAM>
void run(const char* name, int smem, int spin_ns, int grid) {
  auto kern = k<TMEM, BIGPARAM>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  Big a{}, b{};
  int* sink; cudaMalloc(&sink, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 5; ++i) kern<<<grid, 256, smem>>>(a, b, spin_ns, sink);
  cudaDeviceSynchronize();
  const int n = 50;
  cudaEventRecord(e0);
  for (int i = 0; i < n; ++i) kern<<<grid, 256, smem>>>(a, b, spin_ns, sink);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-40s smem=%6d spin=%5.1fus grid=%d : %7.2f us/launch (overhead %6.2f us) err=%s\n", name, smem, spin_ns / 1e3, grid, ms * 1e3 / n, ms * 1e3 / n - spin_ns / 1e3, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  for (int spin : {0, 20000}) {
    run<false, false>("plain", 0, spin, 148);
    run<false, false>("plain 100KB smem", 100 * 1024, spin, 148);
    run<false, false>("plain 227KB smem", 231000, spin, 148);
    run<true, false>("tmem alloc", 0, spin, 148);
    run<true, false>("tmem alloc + 227KB smem", 231000, spin, 148);
    run<false, false>("plain 227KB smem grid 1480", 231000, spin, 1480);
  }
  return 0;
}
