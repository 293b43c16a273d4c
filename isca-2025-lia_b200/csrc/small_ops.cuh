// Device-side bodies of the small HBM-bound kernels (LayerNorm, embedding lookup, greedy argmax), shared by their
// stand-alone kernels (layernorm.cu, misc.cu) and by the decode program kernel (decode_program_sm100.cu), which runs them
// as phases of one persistent launch.  Same arithmetic in the same order in both places: results are bit-identical.
#pragma once
#include "common.cuh"

namespace {

// LayerNorm of rows [row0, row0 + 128/TPR) by a 128-thread group (`tid` in [0,128)): a row lives in registers as 128-bit
// vectors, mean first, then the centred sum of squares.  TPR = threads per row (32: warp per row, h <= 2048; 128: group
// per row, h <= 16384).  `red` = 8 floats of shared memory (TPR == 128 only); `sync128` synchronises the 128 threads.
template <int TPR, int VMAX, typename Sync>
__device__ __forceinline__ void layernorm_rows(const bf16* __restrict__ x, const bf16* __restrict__ w, const bf16* __restrict__ b,
                                               bf16* __restrict__ y, int rows, int h, float eps, int row0, int tid, float* red,
                                               Sync sync128) {
  const int row = row0 + tid / TPR;
  const int t = tid % TPR;
  const int nvec = h >> 3;
  const bool active = row < rows;
  const bf16* xr = x + (size_t)row * h;

  uint4 v[VMAX];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < VMAX; ++i) {
    const int idx = t + i * TPR;
    if (active && idx < nvec) {
      v[i] = ldg_act(xr + idx * 8);
      float f[8];
      unpack8(v[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += f[j];
    }
  }
  sum = warp_sum(sum);
  if (TPR > 32) {
    if ((tid & 31) == 0) red[tid >> 5] = sum;
    sync128();
    sum = red[0] + red[1] + red[2] + red[3];
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
    if ((tid & 31) == 0) red[4 + (tid >> 5)] = sq;
    sync128();
    sq = red[4] + red[5] + red[6] + red[7];
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

// hidden[row,:] = embed_tokens[ids[row]] + embed_positions[p + 2] for one row (b*S + s) by a 128-thread group; the learned
// position p follows OPTLearnedPositionalEmbedding.forward (lia/modeling_opt.py:368-378): cumsum of the mask.  `part` =
// 4 long longs of shared memory.  A null table contributes nothing.
template <typename Sync>
__device__ __forceinline__ void embed_row(const int64_t* __restrict__ ids, const bf16* __restrict__ tok, const bf16* __restrict__ pos,
                                          bf16* __restrict__ out, int row, int S, int h, int past_len, int vocab, int max_pos_rows,
                                          const int64_t* __restrict__ mask, int mask_ld, int tid, long long* part, Sync sync128) {
  const int s = row % S;
  long long id = ids[row];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  long long p = past_len + s;
  if (mask != nullptr) {
    const int64_t* mr = mask + (size_t)(row / S) * mask_ld;
    const int t = past_len + s;
    long long c = 0;
    for (int j = tid; j <= t; j += 128) c += mr[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((tid & 31) == 0) part[tid >> 5] = c;
    sync128();
    p = (part[0] + part[1] + part[2] + part[3]) * mr[t] - 1;
  }
  p += 2;
  p = p < 0 ? 0 : (p >= max_pos_rows ? max_pos_rows - 1 : p);
  const bf16* tr = tok != nullptr ? tok + (size_t)id * h : nullptr;
  const bf16* pr = pos != nullptr ? pos + (size_t)p * h : nullptr;
  bf16* o = out + (size_t)row * h;
  for (int i = tid * 8; i < h; i += 128 * 8) {
    if (tr != nullptr && pr != nullptr) {
      float a[8], c[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(tr + i)), a);
      unpack8(__ldg(reinterpret_cast<const uint4*>(pr + i)), c);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += c[j];
      *reinterpret_cast<uint4*>(o + i) = pack8(a);
    } else {
      *reinterpret_cast<uint4*>(o + i) = __ldg(reinterpret_cast<const uint4*>((tr != nullptr ? tr : pr) + i));
    }
  }
}

// next = argmax_v row[v] by a group of NT threads (lowest index on ties, one id optionally suppressed).  `sv` / `si` =
// NT/32 floats / ints of shared memory.  The result does not depend on NT: ties are resolved by index.
template <int NT, typename Sync>
__device__ __forceinline__ void argmax_row(const bf16* __restrict__ row, int64_t* __restrict__ next, int V, int suppress, int tid,
                                           float* sv, int* si, Sync syncNT) {
  float best = -INFINITY;
  int bi = 0x7fffffff;
  const int nvec = V >> 3;
  for (int i = tid; i < nvec; i += NT) {
    float f[8];
    unpack8(ldg_act(row + i * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int idx = i * 8 + j;
      const float v = (idx == suppress) ? -INFINITY : f[j];
      if (v > best || (v == best && idx < bi)) {
        best = v;
        bi = idx;
      }
    }
  }
  for (int idx = nvec * 8 + tid; idx < V; idx += NT) {
    const float v = (idx == suppress) ? -INFINITY : __bfloat162float(*reinterpret_cast<const volatile bf16*>(row + idx));
    if (v > best || (v == best && idx < bi)) {
      best = v;
      bi = idx;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) {
      best = ov;
      bi = oi;
    }
  }
  if ((tid & 31) == 0) {
    sv[tid >> 5] = best;
    si[tid >> 5] = bi;
  }
  syncNT();
  if (tid == 0) {
    for (int w = 1; w < NT / 32; ++w)
      if (sv[w] > best || (sv[w] == best && si[w] < bi)) {
        best = sv[w];
        bi = si[w];
      }
    *next = (bi == 0x7fffffff) ? 0 : bi;
  }
}

}  // namespace
