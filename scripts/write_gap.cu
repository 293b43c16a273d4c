// Does a kernel pay, at its end, for the bytes it wrote?  148 CTAs x 128 threads write `mb` MB with plain / .cg / fenced stores.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
template <int MODE>
__global__ void __launch_bounds__(128) wk(float4* dst, size_t n_vec, unsigned long long* span) {
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (size_t i = (size_t)blockIdx.x * 128 + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * 128) {
    float4 v = make_float4(1.f, 2.f, 3.f, (float)i);
    if (MODE == 0) dst[i] = v;
    if (MODE == 1) __stcg(dst + i, v);
    if (MODE == 2) { dst[i] = v; }
  }
  if (MODE == 2) __threadfence();
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  if (threadIdx.x == 0) { atomicMin(span, t0); atomicMax(span + 1, t1); }
}
template <int MODE>
void run(const char* name, double mb) {
  size_t n_vec = (size_t)(mb * 1e6 / 16);
  float4* dst; cudaMalloc(&dst, 64 << 20);
  unsigned long long* span; cudaMallocManaged(&span, 16);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 5; ++i) wk<MODE><<<148, 128>>>(dst, n_vec, span);
  cudaDeviceSynchronize();
  span[0] = ~0ull; span[1] = 0;
  wk<MODE><<<148, 128>>>(dst, n_vec, span);
  cudaDeviceSynchronize();
  double in_kernel = (span[1] - span[0]) / 1e3;
  const int n = 50;
  cudaEventRecord(e0);
  for (int i = 0; i < n; ++i) wk<MODE><<<148, 128>>>(dst, n_vec, span);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-10s %5.1f MB: %7.2f us/launch, in-CTA span %7.2f us\n", name, mb, ms * 1e3 / n, in_kernel);
  cudaFree(dst);
}
int main() {
  for (double mb : {0.1, 1.0, 4.0, 8.0, 16.0, 32.0}) { run<0>("plain", mb); run<1>("stcg", mb); run<2>("fenced", mb); }
  return 0;
}
