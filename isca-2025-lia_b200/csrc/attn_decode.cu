// Decode attention (one query token per sequence) over the time-major KV cache.
// Replaces attentions.py:395-399 (cat + permute + contiguous of the whole cache), 493-536
// (3 transposes, bmm, softmax, bmm) for tgt_len == 1.  No mask is applied (attentions.py:500).
//
// HBM-bound: every cached K and V element is read exactly once, in place, with 128-bit
// streaming loads (a (t,b,h) row is d*2 = 128/256 contiguous bytes); the reference moves
// ~5x those bytes.  Flash-decoding: grid = (B*H, splits); a CTA owns keys [t0,t1) of one
// (b,h).  Scores are staged in shared memory, softmax statistics use warp shuffles.
//
// Numerics mirror the reference when splits == 1 (the B=64 bench case):
//   s = bf16(q.k)  [bmm output], p = bf16(exp(s-max)/sum) [softmax(dtype=bf16)], ctx = bf16(sum_t p*v).
// With splits > 1 (small batches) the partial results are combined flash-decoding style
// in fp32 and p is not rounded per element (difference <= 1 bf16 ulp of p).
#include "common.cuh"

namespace {

constexpr int THREADS = 128;
constexpr int MAX_CHUNK = 4096;   // keys per CTA (scores kept in smem)

template <int D>
__global__ void __launch_bounds__(THREADS) attn_decode_kernel(const bf16* __restrict__ q, const bf16* __restrict__ kc,
                                                              const bf16* __restrict__ vc, bf16* __restrict__ out,
                                                              float* __restrict__ ws, int T, int H, int cache_batch,
                                                              int b0, int t_chunk, int splits) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int LPR = D / 8;        // lanes per cached row (each lane owns 8 dims = 16 bytes)
  constexpr int RPW = 32 / LPR;     // rows per warp iteration
  constexpr int UNROLL = 4;
  extern __shared__ float sc[];     // [t_chunk] scores -> probabilities
  __shared__ float red[4];
  __shared__ float accs[4][D];

  const int bh = blockIdx.x;
  const int b = bh / H;
  const int hh = bh - b * H;
  const int split = blockIdx.y;
  const int t0 = split * t_chunk;
  const int n = min(T, t0 + t_chunk) - t0;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int chunk = lane % LPR;
  const int rsub = lane / LPR;

  float qf[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(q + (size_t)bh * D + chunk * 8)), qf);
  const size_t row_stride = (size_t)cache_batch * H * D;
  const size_t base = ((size_t)(b0 + b) * H + hh) * D + chunk * 8 + (size_t)t0 * row_stride;

  // ---- phase A: s_t = bf16(q . k_t)
  for (int i = 0; i < n; i += 4 * RPW * UNROLL) {
    uint4 kv[UNROLL];
    int r[UNROLL];
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      r[j] = i + (j * 4 + warp) * RPW + rsub;
      if (r[j] < n) kv[j] = ldg_stream(kc + base + (size_t)r[j] * row_stride);
    }
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      float dot = 0.f;
      if (r[j] < n) {
        float kf[8];
        unpack8(kv[j], kf);
#pragma unroll
        for (int e = 0; e < 8; ++e) dot = fmaf(qf[e], kf[e], dot);
      }
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      if (chunk == 0 && r[j] < n) sc[r[j]] = bf16r(dot);
    }
  }
  __syncthreads();

  // ---- softmax statistics over the chunk
  float m = -INFINITY;
  for (int i = threadIdx.x; i < n; i += THREADS) m = fmaxf(m, sc[i]);
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float l = 0.f;
  for (int i = threadIdx.x; i < n; i += THREADS) {
    const float e = expf(sc[i] - m);
    sc[i] = e;
    l += e;
  }
  l = warp_sum(l);
  if (lane == 0) red[warp] = l;
  __syncthreads();
  l = red[0] + red[1] + red[2] + red[3];
  if (splits == 1) {
    for (int i = threadIdx.x; i < n; i += THREADS) sc[i] = bf16r(sc[i] / l);
  }
  __syncthreads();

  // ---- phase B: ctx = sum_t p_t * v_t
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  for (int i = 0; i < n; i += 4 * RPW * UNROLL) {
    uint4 vv[UNROLL];
    int r[UNROLL];
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      r[j] = i + (j * 4 + warp) * RPW + rsub;
      if (r[j] < n) vv[j] = ldg_stream(vc + base + (size_t)r[j] * row_stride);
    }
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      if (r[j] < n) {
        const float pj = sc[r[j]];
        float vf[8];
        unpack8(vv[j], vf);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(pj, vf[e], acc[e]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
#pragma unroll
    for (int o = 16; o >= LPR; o >>= 1) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], o);
  }
  if (rsub == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) accs[warp][chunk * 8 + e] = acc[e];
  }
  __syncthreads();
  if (threadIdx.x < D) {
    const int e = threadIdx.x;
    const float o = accs[0][e] + accs[1][e] + accs[2][e] + accs[3][e];
    if (splits == 1) {
      out[(size_t)bh * D + e] = __float2bfloat16_rn(o);
    } else {
      float* dst = ws + ((size_t)bh * splits + split) * (D + 2);
      dst[2 + e] = o;
      if (e == 0) {
        dst[0] = m;
        dst[1] = l;
      }
    }
  }
}

template <int D>
__global__ void attn_decode_combine_kernel(const float* __restrict__ ws, bf16* __restrict__ out, int splits) {
  pdl_launch_dependents();
  pdl_wait();
  const int bh = blockIdx.x;
  const int e = threadIdx.x;
  const float* src = ws + (size_t)bh * splits * (D + 2);
  float M = -INFINITY;
  for (int s = 0; s < splits; ++s) M = fmaxf(M, src[s * (D + 2)]);
  float L = 0.f, o = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float sc = expf(src[s * (D + 2)] - M);
    L += src[s * (D + 2) + 1] * sc;
    o += src[s * (D + 2) + 2 + e] * sc;
  }
  out[(size_t)bh * D + e] = __float2bfloat16_rn(o / L);
}

// CTAs per SM the split choice aims for.  Default 2: the unsharded shapes (B * H in the thousands) keep one CTA per (b, h)
// and with it the reference's per-element rounding of p.  A caller whose rank keeps only H / world heads asks for more
// (`splits` = -6 from the tensor-parallel model): at B * H = 448 (OPT-30B, TP8) one CTA per (b, h) left 3 latency-bound CTAs
// per SM -- 24 us for 9.5 us of K/V bytes -- and two key ranges per (b, h) took a 16-layer stack at TP8 shapes from 152.5 to
// 140.0 us per layer (profiles/README.md).  LIA_ATTN_CTAS_PER_SM overrides the default of the automatic choice (A/B runs).
int pick_splits(int B, int H, int T, int per_sm) {
  int splits = (T + MAX_CHUNK - 1) / MAX_CHUNK;
  if (per_sm <= 0) {
    static int dflt = 0;
    if (dflt == 0) {
      const char* e = getenv("LIA_ATTN_CTAS_PER_SM");
      dflt = (e && atoi(e) > 0) ? atoi(e) : 2;
    }
    per_sm = dflt;
  }
  const int want = (per_sm * lia_sm_count() + B * H - 1) / (B * H);
  const int cap = T / 128 > 1 ? T / 128 : 1;                     // keep >= 128 keys per CTA
  int s = want < cap ? want : cap;
  if (s > 32) s = 32;
  return s > splits ? s : splits;
}

}  // namespace

extern "C" size_t lia_attn_decode_workspace_bytes(int B, int H, int d, int max_splits) {
  if (B <= 0 || H <= 0 || d <= 0) return 0;
  if (max_splits <= 0) max_splits = 32;
  return (size_t)B * H * max_splits * (d + 2) * sizeof(float);
}

extern "C" int lia_attn_decode_bf16(const void* q, const void* k_cache, const void* v_cache, void* out, int B, int H,
                                    int T, int d, int cache_batch, int b0, int splits, void* workspace,
                                    size_t workspace_bytes, lia_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LIA_CHECK_ARG(q && k_cache && v_cache && out, "lia_attn_decode_bf16: null pointer");
  LIA_CHECK_ARG(d == 64 || d == 128, "lia_attn_decode_bf16: head_dim must be 64 or 128 (got %d)", d);
  LIA_CHECK_ARG(B > 0 && H > 0 && T > 0, "lia_attn_decode_bf16: B,H,T must be positive");
  LIA_CHECK_ARG(b0 >= 0 && b0 + B <= cache_batch, "lia_attn_decode_bf16: batch window [%d,%d) outside cache batch %d", b0, b0 + B, cache_batch);
  if (splits <= 0) splits = pick_splits(B, H, T, -splits);   // 0: automatic; -n: automatic, aiming for n CTAs per SM
  if (splits > 1 && (workspace == nullptr || workspace_bytes < (size_t)B * H * splits * (d + 2) * sizeof(float))) {
    splits = (T + MAX_CHUNK - 1) / MAX_CHUNK;   // fall back to the fewest splits that fit shared memory
    LIA_CHECK_ARG(splits == 1, "lia_attn_decode_bf16: T=%d needs a workspace of lia_attn_decode_workspace_bytes()", T);
  }
  int t_chunk = (T + splits - 1) / splits;
  splits = (T + t_chunk - 1) / t_chunk;        // no empty splits
  LIA_CHECK_ARG(t_chunk <= MAX_CHUNK, "lia_attn_decode_bf16: chunk too large");
  const dim3 grid(B * H, splits);
  const size_t smem = (size_t)t_chunk * sizeof(float);
  const bf16* qp = reinterpret_cast<const bf16*>(q);
  const bf16* kp = reinterpret_cast<const bf16*>(k_cache);
  const bf16* vp = reinterpret_cast<const bf16*>(v_cache);
  bf16* op = reinterpret_cast<bf16*>(out);
  float* wsp = reinterpret_cast<float*>(workspace);
  if (d == 128) {
    lia_launch(attn_decode_kernel<128>, dim3(grid), dim3(THREADS), smem, stream, qp, kp, vp, op, wsp, T, H, cache_batch, b0, t_chunk, splits);
    LIA_LAUNCH_CHECK();
    if (splits > 1) lia_launch(attn_decode_combine_kernel<128>, dim3(B * H), dim3(128), 0, stream, wsp, op, splits);
  } else {
    lia_launch(attn_decode_kernel<64>, dim3(grid), dim3(THREADS), smem, stream, qp, kp, vp, op, wsp, T, H, cache_batch, b0, t_chunk, splits);
    LIA_LAUNCH_CHECK();
    if (splits > 1) lia_launch(attn_decode_combine_kernel<64>, dim3(B * H), dim3(64), 0, stream, wsp, op, splits);
  }
  LIA_LAUNCH_CHECK();
  return LIA_OK;
}
