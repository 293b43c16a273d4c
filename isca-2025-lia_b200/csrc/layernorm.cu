// LayerNorm over the last dimension: bf16 in, fp32 statistics, bf16 out.
// Replaces F.layer_norm call sites decoder.py:107-119 and the final norm lia/modeling_opt.py:1563-1564.
//
// HBM-bound (2 bytes read + 2 bytes written per element).  A row lives entirely in
// registers as 128-bit vectors: one pass over global memory, mean first and then the
// centred sum of squares (no E[x^2]-E[x]^2 cancellation).  A row is owned by one warp
// (h <= 2048) or by one 128-thread CTA (h <= 16384), so reductions are warp shuffles plus at
// most one shared-memory hop.
#include "common.cuh"
#include "small_ops.cuh"

namespace {

template <int TPR, int VMAX>
__global__ void __launch_bounds__(128) layernorm_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                        const bf16* __restrict__ b, bf16* __restrict__ y, int rows,
                                                        int h, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int ROWS_PER_CTA = 128 / TPR;
  __shared__ float red[8];
  layernorm_rows<TPR, VMAX>(x, w, b, y, rows, h, eps, blockIdx.x * ROWS_PER_CTA, threadIdx.x, red, [] { __syncthreads(); });
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
