// Causal self-attention over the prompt (prefill) on the 5th-generation tensor cores: tcgen05.mma with the score and
// output accumulators in TMEM, Q / K / V tiles brought in by TMA straight from the activation buffer and the time-major
// KV cache the QKV GEMM epilogue just filled.
//
// Replaces attentions.py:444-449 (mask rebuild), 493-496 (3 transposes), 499-512 (bmm, mask add, clamp, softmax over a
// materialised [B,H,S,S] score tensor), 529 (bmm) -- and this library's own first-generation mma.sync kernel
// (attn_prefill.cu), which stays as the A/B baseline (LIA_ATTN_PREFILL_TC=0).
//
// Work item = (sequence b, head h, 128-query tile); persistent CTAs (one per SM) walk the item list heaviest-first.
// Per item, TWO passes over the causal key blocks (128 keys each), because the reference's rounding points need the
// exact row maximum m and sum l before any probability is formed:
//   pass 1:  S_j = Q K_j^T (tcgen05, fp32 in TMEM) -> s = bf16(S) (the bmm's output dtype), causal mask, m, l
//   pass 2:  S_j again, p = bf16(exp(s - m) / l)   (= softmax(dtype=bf16)), P_j -> shared memory (K-major, 128-byte
//            swizzle) -> O += P_j V_j (tcgen05; V_j is the MN-major operand, exactly as TMA lays the cache rows down)
//   ctx = bf16(O)
// Recomputing Q K^T costs tensor time that is idle anyway: the kernel is bound by the softmax arithmetic (one ex2 per
// score and pass on the 16-lane MUFU pipe), not by the MMAs.
//
// Warp roles (320 threads): warps 0-7 softmax + epilogue -- a thread owns ONE query row (= one TMEM lane), so row max
//   and row sum need no shuffles; warps w and w+4 share a lane quarter and split each score block by columns
//   (keys 0-63 | 64-127), which doubles the issue slots available to the softmax; their partial (m, l) meet in shared
//   memory once per item.  warp 8: TMA producer (Q once per item, K tiles for both passes, V tiles for pass 2; 2-stage
//   rings).  warp 9: TMEM allocator + MMA issuer (one lane): scores are double-buffered in TMEM (2 x 128 columns) so
//   Q K_{j+1}^T runs while the softmax warps work on block j; O takes `d` more columns.
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace {

constexpr int BM = 128;              // queries per tile = TMEM lanes = UMMA M
constexpr int BN = 128;              // keys per block
constexpr int SOFTMAX_THREADS = 256;
constexpr int PRODUCER_WARP = 8;
constexpr int MMA_WARP = 9;
constexpr int THREADS = 320;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float M_INIT = -1.0e30f;   // finite "no key seen yet": keeps (m_old - m_new) and (s - m) free of inf - inf

template <int D>
struct Layout {
  static constexpr int SLABS = D / 64;               // 64-element (128-byte) wide column slabs per row
  static constexpr int SLAB_BYTES = BM * 128;        // 128 rows x 128 bytes
  static constexpr int TILE_BYTES = SLABS * SLAB_BYTES;
  static constexpr int Q_OFF = 0;
  static constexpr int K_OFF = Q_OFF + TILE_BYTES;           // 2 stages
  static constexpr int V_OFF = K_OFF + 2 * TILE_BYTES;       // 2 stages
  static constexpr int P_OFF = V_OFF + 2 * TILE_BYTES;       // [128 queries][128 keys] bf16 = 2 slabs
  static constexpr int STAT_OFF = P_OFF + 2 * SLAB_BYTES;    // float2 [2 item parities][2 column halves][128 rows]
  static constexpr int BAR_OFF = STAT_OFF + 2 * 2 * BM * 8;
  static constexpr int NUM_BARS = 18;
  static constexpr int TOTAL = BAR_OFF + NUM_BARS * 8 + 16 + 1024 /* alignment slack */;
  static constexpr int O_COL = 2 * BN;                       // TMEM: S[0] | S[1] | O
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void softmax_bar_sync() { asm volatile("bar.sync 2, 256;" ::: "memory"); }

struct Item {
  int b, hh, qt, nblk;
};
__device__ __forceinline__ Item decode_item(int idx, int bh_count, int H, int nq) {
  Item it;
  it.qt = nq - 1 - idx / bh_count;            // heaviest (most key blocks) first
  const int bh = idx - (idx / bh_count) * bh_count;
  it.b = bh / H;
  it.hh = bh - it.b * H;
  it.nblk = it.qt + 1;                        // causal: key blocks 0..qt
  return it;
}

template <int D>
__global__ void __launch_bounds__(THREADS, 1)
attn_prefill_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, bf16* __restrict__ out, int H, int S, int b0, int nq,
                       int n_items, int bh_count) {
  using L = Layout<D>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + L::BAR_OFF;
  // barrier map
  const uint32_t q_full = bar_base, q_empty = bar_base + 8;
  auto k_full = [&](int s) { return bar_base + 16 + 8u * s; };
  auto k_empty = [&](int s) { return bar_base + 32 + 8u * s; };
  auto v_full = [&](int s) { return bar_base + 48 + 8u * s; };
  auto v_empty = [&](int s) { return bar_base + 64 + 8u * s; };
  auto s_full = [&](int s) { return bar_base + 80 + 8u * s; };
  auto s_empty = [&](int s) { return bar_base + 96 + 8u * s; };
  const uint32_t p_full = bar_base + 112, p_empty = bar_base + 120, o_full = bar_base + 128, o_empty = bar_base + 136;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + L::BAR_OFF + L::NUM_BARS * 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == PRODUCER_WARP && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(k_full(s), 1);
      mbar_init(k_empty(s), 1);
      mbar_init(v_full(s), 1);
      mbar_init(v_empty(s), 1);
      mbar_init(s_full(s), 1);
      mbar_init(s_empty(s), 8);        // one arrival per softmax warp
    }
    mbar_init(p_full, 8);
    mbar_init(p_empty, 1);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr_smem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();

  if (warp == PRODUCER_WARP) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      pdl_wait();                                   // q and the cache rows come from the QKV GEMM before us
      unsigned kcount = 0, vcount = 0, n = 0;
      auto load_tile = [&](const CUtensorMap* map, uint32_t dst, int col0, int row0, uint32_t bar) {
        mbar_expect_tx(bar, L::TILE_BYTES);
#pragma unroll
        for (int sl = 0; sl < L::SLABS; ++sl) tma_load_2d(dst + sl * L::SLAB_BYTES, map, col0 + sl * 64, row0, bar);
      };
      auto load_k = [&](const Item& it, int j) {
        const int st = kcount & 1;
        mbar_wait(k_empty(st), ((kcount >> 1) & 1) ^ 1u);
        load_tile(&tmK, smem_base + L::K_OFF + st * L::TILE_BYTES, ((b0 + it.b) * H + it.hh) * D, j * BN, k_full(st));
        ++kcount;
      };
      for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x, ++n) {
        const Item it = decode_item(idx, bh_count, H, nq);
        mbar_wait(q_empty, (n & 1) ^ 1u);
        load_tile(&tmQ, smem_base + L::Q_OFF, it.hh * D, it.b * S + it.qt * BM, q_full);
        for (int j = 0; j < it.nblk; ++j) load_k(it, j);                 // pass 1
        for (int j = 0; j < it.nblk; ++j) {                              // pass 2
          load_k(it, j);
          const int st = vcount & 1;
          mbar_wait(v_empty(st), ((vcount >> 1) & 1) ^ 1u);
          load_tile(&tmV, smem_base + L::V_OFF + st * L::TILE_BYTES, ((b0 + it.b) * H + it.hh) * D, j * BN, v_full(st));
          ++vcount;
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc(BM, BN);                   // A = Q (K-major), B = K (K-major)
      constexpr uint32_t idesc_pv = make_idesc(BM, D) | (1u << 16);       // A = P (K-major), B = V (MN-major)
      unsigned kc = 0, vc = 0, sc = 0, pc = 0, n = 0;
      auto issue_qk = [&]() {
        const int ks = kc & 1, sb = sc & 1;
        mbar_wait(k_full(ks), (kc >> 1) & 1);
        mbar_wait(s_empty(sb), ((sc >> 1) & 1) ^ 1u);
        tcgen05_fence_after();
        const uint32_t qa = smem_base + L::Q_OFF, ka = smem_base + L::K_OFF + ks * L::TILE_BYTES;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t off = (kk >> 2) * L::SLAB_BYTES + (kk & 3) * 32;   // 64-element slab, 16 elements = 32 bytes inside the swizzle row
          tcgen05_mma_f16(tmem_base + sb * BN, make_smem_desc(qa + off), make_smem_desc(ka + off), idesc_qk, kk > 0 ? 1u : 0u);
        }
        tcgen05_commit(k_empty(ks));
        tcgen05_commit(s_full(sb));
        ++kc;
        ++sc;
      };
      for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x, ++n) {
        const Item it = decode_item(idx, bh_count, H, nq);
        mbar_wait(q_full, n & 1);
        for (int j = 0; j < it.nblk; ++j) issue_qk();                    // pass 1: scores only
        issue_qk();                                                      // pass 2, block 0
        for (int j = 0; j < it.nblk; ++j) {
          if (j + 1 < it.nblk) issue_qk();                               // next block's scores while the softmax works on this one
          const int vs = vc & 1;
          mbar_wait(v_full(vs), (vc >> 1) & 1);
          mbar_wait(p_full, pc & 1);
          if (j == 0) mbar_wait(o_empty, (n & 1) ^ 1u);                  // the previous item's O has been read out
          tcgen05_fence_after();
          const uint32_t pa = smem_base + L::P_OFF, va = smem_base + L::V_OFF + vs * L::TILE_BYTES;
#pragma unroll
          for (int kk = 0; kk < BN / 16; ++kk) {
            const uint64_t dp = make_smem_desc(pa + (kk >> 2) * L::SLAB_BYTES + (kk & 3) * 32);
            const uint64_t dv = make_smem_desc_mn(va + kk * 16 * 128, L::SLAB_BYTES);   // 16 keys = 16 rows of 128 bytes
            tcgen05_mma_f16(tmem_base + L::O_COL, dp, dv, idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
          }
          tcgen05_commit(v_empty(vs));
          tcgen05_commit(p_empty);
          ++vc;
          ++pc;
        }
        tcgen05_commit(o_full);
        tcgen05_commit(q_empty);
      }
    }
  } else {
    // ===================== softmax + epilogue (256 threads) =====================
    const int g = warp >> 2;                       // column half of every score block / of O
    const int wq = warp & 3;                       // TMEM lane quarter
    const int row = wq * 32 + lane;                // query row inside the tile = TMEM lane
    const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
    float2* stat = reinterpret_cast<float2*>(smem_gen + L::STAT_OFF);
    unsigned sc = 0, pcnt = 0, n = 0;
    for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x, ++n) {
      const Item it = decode_item(idx, bh_count, H, nq);
      // ---------------- pass 1: m and l over this thread's 64 columns of every block
      float m = M_INIT, l = 0.f;
      for (int j = 0; j < it.nblk; ++j) {
        const int sb = sc & 1;
        mbar_wait(s_full(sb), (sc >> 1) & 1);
        tcgen05_fence_after();
        uint32_t v[64];
        const uint32_t taddr = tmem_base + lane_addr + sb * BN + g * 64;
        tmem_ld32(taddr, v);
        tmem_ld32(taddr + 32, v + 32);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty(sb));
        ++sc;
        const bool diag = (j == it.qt);
        float bm = -INFINITY;
#pragma unroll
        for (int i = 0; i < 64; i += 2) {
          float s0, s1;
          unpack_bf16x2(pack_bf16x2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), s0, s1);   // s = bf16(acc): the bmm's output
          if (diag) {
            if (g * 64 + i > row) s0 = -INFINITY;
            if (g * 64 + i + 1 > row) s1 = -INFINITY;
          }
          v[i] = __float_as_uint(s0);
          v[i + 1] = __float_as_uint(s1);
          bm = fmaxf(bm, fmaxf(s0, s1));
        }
        const float mn = fmaxf(m, bm);
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 64; ++i) acc += ex2((__uint_as_float(v[i]) - mn) * LOG2E);
        l = l * ex2((m - mn) * LOG2E) + acc;
        m = mn;
      }
      // the two column halves of a row meet: m = max, l rescaled to it
      stat[((n & 1) * 2 + g) * BM + row] = make_float2(m, l);
      softmax_bar_sync();
      const float2 o2 = stat[((n & 1) * 2 + (g ^ 1)) * BM + row];
      const float mf = fmaxf(m, o2.x);
      const float lf = l * ex2((m - mf) * LOG2E) + o2.y * ex2((o2.x - mf) * LOG2E);
      const float inv_l = 1.f / lf;
      // ---------------- pass 2: p = bf16(exp(s - m) / l) -> shared memory (this half's 64 keys = one K-major slab)
      for (int j = 0; j < it.nblk; ++j) {
        const int sb = sc & 1;
        mbar_wait(s_full(sb), (sc >> 1) & 1);
        tcgen05_fence_after();
        uint32_t v[64];
        const uint32_t taddr = tmem_base + lane_addr + sb * BN + g * 64;
        tmem_ld32(taddr, v);
        tmem_ld32(taddr + 32, v + 32);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty(sb));
        ++sc;
        const bool diag = (j == it.qt);
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 64; i += 2) {
          float s0, s1;
          unpack_bf16x2(pack_bf16x2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), s0, s1);
          float p0 = ex2((s0 - mf) * LOG2E) * inv_l;
          float p1 = ex2((s1 - mf) * LOG2E) * inv_l;
          if (diag) {
            if (g * 64 + i > row) p0 = 0.f;
            if (g * 64 + i + 1 > row) p1 = 0.f;
          }
          pk[i >> 1] = pack_bf16x2(p0, p1);
        }
        mbar_wait(p_empty, (pcnt & 1) ^ 1u);          // the previous block's P.V has finished reading the P tile
        const uint32_t prow = smem_base + L::P_OFF + g * L::SLAB_BYTES + row * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t addr = prow + ((c ^ (row & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(pk[4 * c]), "r"(pk[4 * c + 1]), "r"(pk[4 * c + 2]),
                       "r"(pk[4 * c + 3])
                       : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        ++pcnt;
      }
      // ---------------- epilogue: ctx = bf16(O), this half's d/2 columns
      mbar_wait(o_full, n & 1);
      tcgen05_fence_after();
      constexpr int OC = D / 2;
      uint32_t o[OC];
      const uint32_t oaddr = tmem_base + lane_addr + L::O_COL + g * OC;
      tmem_ld32(oaddr, o);
      if constexpr (OC == 64) tmem_ld32(oaddr + 32, o + 32);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      const int qrow = it.qt * BM + row;
      if (qrow < S) {
        bf16* dst = out + ((size_t)(it.b * S + qrow) * H + it.hh) * D + g * OC;
#pragma unroll
        for (int c = 0; c < OC / 8; ++c) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[8 * c + 0]), __uint_as_float(o[8 * c + 1]));
          u.y = pack_bf16x2(__uint_as_float(o[8 * c + 2]), __uint_as_float(o[8 * c + 3]));
          u.z = pack_bf16x2(__uint_as_float(o[8 * c + 4]), __uint_as_float(o[8 * c + 5]));
          u.w = pack_bf16x2(__uint_as_float(o[8 * c + 6]), __uint_as_float(o[8 * c + 7]));
          *reinterpret_cast<uint4*>(dst + 8 * c) = u;
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <int D>
int launch_tc(const bf16* q, const bf16* kc, const bf16* vc, bf16* out, int B, int H, int S, int cache_batch, int b0, int t_rows,
              cudaStream_t stream) {
  using L = Layout<D>;
  static_assert(L::TOTAL <= 232448, "shared memory budget exceeded");
  auto kern = attn_prefill_tc_kernel<D>;
  static bool configured = false;
  if (!configured) {
    LIA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  // q: [B*S rows, H*d]; caches: [t_rows, cache_batch*H*d] (time-major rows: a (b,h) slice is a column window)
  if ((rc = lia_make_tmap_2d(&tmQ, q, (uint64_t)B * S, (uint64_t)H * D, (uint64_t)H * D, BM)) != LIA_OK) return rc;
  if ((rc = lia_make_tmap_2d(&tmK, kc, (uint64_t)t_rows, (uint64_t)cache_batch * H * D, (uint64_t)cache_batch * H * D, BN)) != LIA_OK) return rc;
  if ((rc = lia_make_tmap_2d(&tmV, vc, (uint64_t)t_rows, (uint64_t)cache_batch * H * D, (uint64_t)cache_batch * H * D, BN)) != LIA_OK) return rc;
  const int nq = (S + BM - 1) / BM;
  const long long items = (long long)B * H * nq;
  LIA_CHECK_ARG(items < (1ll << 31), "lia_attn_prefill_bf16: too many work items");
  int grid = lia_sm_count();
  if (items < grid) grid = (int)items;
  LIA_CUDA(lia_launch(kern, dim3(grid), dim3(THREADS), L::TOTAL, stream, tmQ, tmK, tmV, out, H, S, b0, nq, (int)items, B * H));
  return LIA_OK;
}

}  // namespace

// `t_rows` = rows the cache tensors really have (>= S): the TMA tensor map must not describe memory past the allocation
int lia_attn_prefill_tc(const void* q, const void* k_cache, const void* v_cache, void* out, int B, int H, int S, int d,
                        int cache_batch, int b0, int t_rows, cudaStream_t stream) {
  const bf16* qp = reinterpret_cast<const bf16*>(q);
  const bf16* kp = reinterpret_cast<const bf16*>(k_cache);
  const bf16* vp = reinterpret_cast<const bf16*>(v_cache);
  bf16* op = reinterpret_cast<bf16*>(out);
  if (d == 128) return launch_tc<128>(qp, kp, vp, op, B, H, S, cache_batch, b0, t_rows, stream);
  return launch_tc<64>(qp, kp, vp, op, B, H, S, cache_batch, b0, t_rows, stream);
}
