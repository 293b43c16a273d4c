// LayerNorm over the last dimension: bf16 in, fp32 statistics, bf16 out.
// Replaces F.layer_norm call sites decoder.py:107-119 and the final norm lia/modeling_opt.py:1563-1564.
//
// HBM-bound (2 bytes read + 2 bytes written per element).  A row lives entirely in
// registers as 128-bit vectors: one pass over global memory, mean first and then the
// centred sum of squares (no E[x^2]-E[x]^2 cancellation).  A row is owned by one warp
// (h <= 2048) or by one 128-thread CTA (h <= 16384), so reductions are warp shuffles plus at
// most one shared-memory hop.
#include "common.cuh"

namespace {

template <int TPR, int VMAX>
__global__ void __launch_bounds__(128) layernorm_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                        const bf16* __restrict__ b, bf16* __restrict__ y, int rows,
                                                        int h, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int ROWS_PER_CTA = 128 / TPR;
  __shared__ float red[2][4];
  const int row = blockIdx.x * ROWS_PER_CTA + threadIdx.x / TPR;
  const int t = threadIdx.x % TPR;
  const int nvec = h >> 3;
  const bool active = row < rows;
  const bf16* xr = x + (size_t)row * h;

  uint4 v[VMAX];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < VMAX; ++i) {
    const int idx = t + i * TPR;
    if (active && idx < nvec) {
      v[i] = ldg_stream(xr + idx * 8);
      float f[8];
      unpack8(v[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += f[j];
    }
  }
  sum = warp_sum(sum);
  if (TPR > 32) {
    if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = sum;
    __syncthreads();
    sum = red[0][0] + red[0][1] + red[0][2] + red[0][3];
  }
  const float mean = sum / (float)h;

  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < VMAX; ++i) {
    const int idx = t + i * TPR;
    if (active && idx < nvec) {
      float f[8];
      unpack8(v[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dlt = f[j] - mean;
        sq += dlt * dlt;
      }
    }
  }
  sq = warp_sum(sq);
  if (TPR > 32) {
    if ((threadIdx.x & 31) == 0) red[1][threadIdx.x >> 5] = sq;
    __syncthreads();
    sq = red[1][0] + red[1][1] + red[1][2] + red[1][3];
  }
  const float rstd = rsqrtf(sq / (float)h + eps);

  bf16* yr = y + (size_t)row * h;
#pragma unroll
  for (int i = 0; i < VMAX; ++i) {
    const int idx = t + i * TPR;
    if (active && idx < nvec) {
      float f[8], g[8], bb[8];
      unpack8(v[i], f);
      unpack8(__ldg(reinterpret_cast<const uint4*>(w + idx * 8)), g);
      unpack8(__ldg(reinterpret_cast<const uint4*>(b + idx * 8)), bb);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (f[j] - mean) * rstd * g[j] + bb[j];
      *reinterpret_cast<uint4*>(yr + idx * 8) = pack8(f);
    }
  }
}

}  // namespace

extern "C" int lia_layernorm_bf16(const void* x, const void* w, const void* b, void* y, int rows, int h, float eps,
                                  lia_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LIA_CHECK_ARG(x && w && b && y, "lia_layernorm_bf16: null pointer");
  LIA_CHECK_ARG(rows >= 0 && h > 0 && h % 8 == 0 && h <= 16384, "lia_layernorm_bf16: need h %% 8 == 0 and h <= 16384 (h=%d)", h);
  if (rows == 0) return LIA_OK;
  const bf16* xp = reinterpret_cast<const bf16*>(x);
  const bf16* wp = reinterpret_cast<const bf16*>(w);
  const bf16* bp = reinterpret_cast<const bf16*>(b);
  bf16* yp = reinterpret_cast<bf16*>(y);
  if (h <= 2048) {
    lia_launch(layernorm_kernel<32, 8>, dim3((rows + 3) / 4), dim3(128), 0, stream, xp, wp, bp, yp, rows, h, eps);
  } else {
    lia_launch(layernorm_kernel<128, 16>, dim3(rows), dim3(128), 0, stream, xp, wp, bp, yp, rows, h, eps);
  }
  LIA_LAUNCH_CHECK();
  return LIA_OK;
}
