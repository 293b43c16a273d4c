// Causal self-attention over the prompt (prefill), reading K/V straight from the time-major
// KV cache the QKV GEMM epilogue just filled.
// Replaces attentions.py:444-449 (mask rebuild), 493-496 (3 transposes), 499-512 (bmm, mask add,
// clamp, softmax with materialised [B,H,S,S] scores), 529 (bmm).
//
// FIRST-GENERATION kernel, kept as the A/B baseline (LIA_ATTN_PREFILL_TC=0) of the tcgen05 kernel in
// attn_prefill_sm100.cu, which is the default: a register-resident flash-style kernel on mma.sync.m16n8k16
// (measured 78 TFLOP/s at B32 H56 S256 d128 -- 5 % of the tensor peak, profiles/r2/r2_ab_attn_prefill.log).
//
// To reproduce the reference's rounding points exactly, softmax is done in TWO passes over the
// keys instead of with online rescaling:
//   pass 1: s = bf16(q.k) (bmm output), causal mask, exact row max m and l = sum exp(s - m)
//   pass 2: p = bf16(exp(s - m) / l)  (softmax(dtype=bf16) output), O += p . v, ctx = bf16(O)
// The extra QK^T costs 0.5x of a tiny kernel and keeps P bit-compatible with the reference.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int TILE = 64;      // queries per CTA and keys per smem tile
constexpr int THREADS = 128;  // 4 warps x 16 query rows

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int D>
__global__ void __launch_bounds__(THREADS) attn_prefill_kernel(const bf16* __restrict__ q, const bf16* __restrict__ kc,
                                                               const bf16* __restrict__ vc, bf16* __restrict__ out,
                                                               int H, int S, int cache_batch, int b0) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int PITCH = D + 8;               // elements; (D+8)*2 bytes keeps ldmatrix rows on distinct banks
  constexpr int TILE_ELEMS = TILE * PITCH;
  constexpr int KSTEPS = D / 16;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sK = reinterpret_cast<bf16*>(smem_raw);            // [2][TILE][PITCH]
  bf16* sV = sK + 2 * TILE_ELEMS;                          // [2][TILE][PITCH]

  const int qi = gridDim.x - 1 - blockIdx.x;               // heaviest (most keys) tiles first
  const int hh = blockIdx.y;
  const int b = blockIdx.z;
  const int q0 = qi * TILE;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2;
  const int tq = lane & 3;
  const int row0 = q0 + warp * 16 + g;
  const int row1 = row0 + 8;
  const size_t kv_row_stride = (size_t)cache_batch * H * D;
  const size_t kv_base = ((size_t)(b0 + b) * H + hh) * D;

  // Q fragments stay in registers for the whole kernel
  uint32_t qf[KSTEPS][4];
  {
    const bf16* q_r0 = q + ((size_t)(b * S + row0) * H + hh) * D;
    const bf16* q_r1 = q + ((size_t)(b * S + row1) * H + hh) * D;
#pragma unroll
    for (int kk = 0; kk < KSTEPS; ++kk) {
      const int c = kk * 16 + tq * 2;
      qf[kk][0] = row0 < S ? *reinterpret_cast<const uint32_t*>(q_r0 + c) : 0u;
      qf[kk][1] = row1 < S ? *reinterpret_cast<const uint32_t*>(q_r1 + c) : 0u;
      qf[kk][2] = row0 < S ? *reinterpret_cast<const uint32_t*>(q_r0 + c + 8) : 0u;
      qf[kk][3] = row1 < S ? *reinterpret_cast<const uint32_t*>(q_r1 + c + 8) : 0u;
    }
  }

  auto load_tile = [&](int stage, int j, bool with_v) {
    constexpr int CPR = D / 8;   // 16-byte chunks per row
    for (int idx = threadIdx.x; idx < TILE * CPR; idx += THREADS) {
      const int r = idx / CPR;
      const int c = idx - r * CPR;
      const int t = j * TILE + r;
      bf16* dk = sK + stage * TILE_ELEMS + r * PITCH + c * 8;
      bf16* dv = sV + stage * TILE_ELEMS + r * PITCH + c * 8;
      if (t < S) {
        const size_t off = (size_t)t * kv_row_stride + kv_base + c * 8;
        cp_async16(smem_u32(dk), kc + off);
        if (with_v) cp_async16(smem_u32(dv), vc + off);
      } else {
        *reinterpret_cast<uint4*>(dk) = make_uint4(0, 0, 0, 0);
        if (with_v) *reinterpret_cast<uint4*>(dv) = make_uint4(0, 0, 0, 0);
      }
    }
    cp_async_commit();
  };

  // s = bf16(q.k) for one 16x64 slab, causal mask applied (masked -> -inf)
  auto score_tile = [&](int stage, int j, float (*s)[4]) {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
    const uint32_t kbase = smem_u32(sK + stage * TILE_ELEMS);
#pragma unroll
    for (int kk = 0; kk < KSTEPS; ++kk) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t r0, r1, r2, r3;
        const int krow = np * 16 + (lane >> 4) * 8 + (lane & 7);
        const int kcol = kk * 16 + ((lane >> 3) & 1) * 8;
        ldmatrix_x4(kbase + (krow * PITCH + kcol) * 2, r0, r1, r2, r3);
        mma_bf16(s[2 * np], qf[kk], r0, r1);
        mma_bf16(s[2 * np + 1], qf[kk], r2, r3);
      }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int key = j * TILE + nt * 8 + tq * 2;
      s[nt][0] = (key <= row0) ? bf16r(s[nt][0]) : -INFINITY;
      s[nt][1] = (key + 1 <= row0) ? bf16r(s[nt][1]) : -INFINITY;
      s[nt][2] = (key <= row1) ? bf16r(s[nt][2]) : -INFINITY;
      s[nt][3] = (key + 1 <= row1) ? bf16r(s[nt][3]) : -INFINITY;
    }
  };

  // ---------------- pass 1: row max and sum of exponentials
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  load_tile(0, 0, false);
  for (int j = 0; j <= qi; ++j) {
    if (j < qi) {
      load_tile((j + 1) & 1, j + 1, false);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    float s[8][4];
    score_tile(j & 1, j, s);
    float t0 = -INFINITY, t1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      t0 = fmaxf(t0, fmaxf(s[nt][0], s[nt][1]));
      t1 = fmaxf(t1, fmaxf(s[nt][2], s[nt][3]));
    }
    t0 = fmaxf(t0, __shfl_xor_sync(0xffffffffu, t0, 1));
    t0 = fmaxf(t0, __shfl_xor_sync(0xffffffffu, t0, 2));
    t1 = fmaxf(t1, __shfl_xor_sync(0xffffffffu, t1, 1));
    t1 = fmaxf(t1, __shfl_xor_sync(0xffffffffu, t1, 2));
    const float n0 = fmaxf(m0, t0), n1 = fmaxf(m1, t1);
    float a0 = 0.f, a1 = 0.f;
    if (n0 > -INFINITY) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) a0 += expf(s[nt][0] - n0) + expf(s[nt][1] - n0);
    }
    if (n1 > -INFINITY) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) a1 += expf(s[nt][2] - n1) + expf(s[nt][3] - n1);
    }
    a0 += __shfl_xor_sync(0xffffffffu, a0, 1);
    a0 += __shfl_xor_sync(0xffffffffu, a0, 2);
    a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
    a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
    l0 = (m0 > -INFINITY ? l0 * expf(m0 - n0) : 0.f) + a0;
    l1 = (m1 > -INFINITY ? l1 * expf(m1 - n1) : 0.f) + a1;
    m0 = n0;
    m1 = n1;
    __syncthreads();
  }
  const float mm0 = m0 > -INFINITY ? m0 : 0.f, mm1 = m1 > -INFINITY ? m1 : 0.f;
  const float ll0 = l0 > 0.f ? l0 : 1.f, ll1 = l1 > 0.f ? l1 : 1.f;

  // ---------------- pass 2: p = bf16(exp(s-m)/l), O += p.v
  float o[D / 8][4];
#pragma unroll
  for (int nt = 0; nt < D / 8; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
  load_tile(0, 0, true);
  for (int j = 0; j <= qi; ++j) {
    if (j < qi) {
      load_tile((j + 1) & 1, j + 1, true);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    float s[8][4];
    score_tile(j & 1, j, s);
    uint32_t pf[4][4];   // P as A-fragments: 4 k-steps of 16 keys
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) {
      pf[k2][0] = pack_bf16x2((expf(s[2 * k2][0] - mm0) / ll0), (expf(s[2 * k2][1] - mm0) / ll0));
      pf[k2][1] = pack_bf16x2((expf(s[2 * k2][2] - mm1) / ll1), (expf(s[2 * k2][3] - mm1) / ll1));
      pf[k2][2] = pack_bf16x2((expf(s[2 * k2 + 1][0] - mm0) / ll0), (expf(s[2 * k2 + 1][1] - mm0) / ll0));
      pf[k2][3] = pack_bf16x2((expf(s[2 * k2 + 1][2] - mm1) / ll1), (expf(s[2 * k2 + 1][3] - mm1) / ll1));
    }
    const uint32_t vbase = smem_u32(sV + (j & 1) * TILE_ELEMS);
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) {
#pragma unroll
      for (int dp = 0; dp < D / 16; ++dp) {
        uint32_t r0, r1, r2, r3;
        const int vrow = k2 * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
        const int vcol = dp * 16 + (lane >> 4) * 8;
        ldmatrix_x4_trans(vbase + (vrow * PITCH + vcol) * 2, r0, r1, r2, r3);
        mma_bf16(o[2 * dp], pf[k2], r0, r1);
        mma_bf16(o[2 * dp + 1], pf[k2], r2, r3);
      }
    }
    __syncthreads();
  }

  bf16* o_r0 = out + ((size_t)(b * S + row0) * H + hh) * D;
  bf16* o_r1 = out + ((size_t)(b * S + row1) * H + hh) * D;
#pragma unroll
  for (int nt = 0; nt < D / 8; ++nt) {
    const int c = nt * 8 + tq * 2;
    if (row0 < S) *reinterpret_cast<uint32_t*>(o_r0 + c) = pack_bf16x2(o[nt][0], o[nt][1]);
    if (row1 < S) *reinterpret_cast<uint32_t*>(o_r1 + c) = pack_bf16x2(o[nt][2], o[nt][3]);
  }
}

template <int D>
int launch_prefill(const bf16* q, const bf16* kc, const bf16* vc, bf16* out, int B, int H, int S, int cache_batch, int b0,
                   cudaStream_t stream) {
  constexpr int SMEM = 2 * 2 * TILE * (D + 8) * 2;
  auto kern = attn_prefill_kernel<D>;
  static bool configured = false;
  if (!configured) {
    LIA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  const dim3 grid((S + TILE - 1) / TILE, H, B);
  LIA_CUDA(lia_launch(kern, dim3(grid), dim3(THREADS), SMEM, stream, q, kc, vc, out, H, S, cache_batch, b0));
  return LIA_OK;
}

// The tcgen05 kernel (attn_prefill_sm100.cu) is the default; LIA_ATTN_PREFILL_TC=0 selects the mma.sync kernel above
// for A/B runs (read per call so that one process can compare the two).
bool tc_enabled() {
  const char* e = getenv("LIA_ATTN_PREFILL_TC");
  return e == nullptr || atoi(e) != 0;
}

}  // namespace

int lia_attn_prefill_tc(const void* q, const void* k_cache, const void* v_cache, void* out, int B, int H, int S, int d,
                        int cache_batch, int b0, int t_rows, cudaStream_t stream);   // attn_prefill_sm100.cu

extern "C" int lia_attn_prefill_bf16(const void* q, const void* k_cache, const void* v_cache, void* out, int B, int H,
                                     int S, int d, int cache_batch, int b0, lia_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LIA_CHECK_ARG(q && k_cache && v_cache && out, "lia_attn_prefill_bf16: null pointer");
  LIA_CHECK_ARG(d == 64 || d == 128, "lia_attn_prefill_bf16: head_dim must be 64 or 128 (got %d)", d);
  LIA_CHECK_ARG(B > 0 && H > 0 && S > 0, "lia_attn_prefill_bf16: B,H,S must be positive");
  LIA_CHECK_ARG(B <= 65535 && H <= 65535, "lia_attn_prefill_bf16: B,H exceed grid limits");
  LIA_CHECK_ARG(b0 >= 0 && b0 + B <= cache_batch, "lia_attn_prefill_bf16: batch window [%d,%d) outside cache batch %d", b0, b0 + B, cache_batch);
  const bool aligned = ((uintptr_t)q % 16 == 0) && ((uintptr_t)k_cache % 16 == 0) && ((uintptr_t)v_cache % 16 == 0) && ((uintptr_t)out % 16 == 0);
  if (tc_enabled() && aligned)   // rows [0, S) of the caches are all the tensor maps describe: later rows read as zeros
    return lia_attn_prefill_tc(q, k_cache, v_cache, out, B, H, S, d, cache_batch, b0, S, stream);
  const bf16* qp = reinterpret_cast<const bf16*>(q);
  const bf16* kp = reinterpret_cast<const bf16*>(k_cache);
  const bf16* vp = reinterpret_cast<const bf16*>(v_cache);
  bf16* op = reinterpret_cast<bf16*>(out);
  if (d == 128) return launch_prefill<128>(qp, kp, vp, op, B, H, S, cache_batch, b0, stream);
  return launch_prefill<64>(qp, kp, vp, op, B, H, S, cache_batch, b0, stream);
}
