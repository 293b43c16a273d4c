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
constexpr int QK_PRODUCER_WARP = 8;
constexpr int MMA_WARP = 9;
constexpr int V_PRODUCER_WARP = 10;
constexpr int THREADS = 352;
constexpr int S_BUFS = 3;            // score buffers in TMEM (128 columns each)
constexpr int RESIDENT_BLOCKS = 2;   // items with at most this many key blocks keep their scores in TMEM between the passes
constexpr float LOG2E = 1.4426950408889634f;
constexpr float M_INIT = -1.0e30f;   // finite "no key seen yet": keeps (m_old - m_new) and (s - m) free of inf - inf

template <int D>
struct Layout {
  static constexpr int SLABS = D / 64;               // 64-element (128-byte) wide column slabs per row
  static constexpr int SLAB_BYTES = BM * 128;        // 128 rows x 128 bytes
  static constexpr int TILE_BYTES = SLABS * SLAB_BYTES;
  static constexpr int Q_OFF = 0;                            // 2 buffers (item parity)
  static constexpr int K_OFF = Q_OFF + 2 * TILE_BYTES;       // 2 stages
  static constexpr int V_OFF = K_OFF + 2 * TILE_BYTES;       // 1 stage
  static constexpr int P_OFF = V_OFF + TILE_BYTES;           // [128 queries][128 keys] bf16 = 2 slabs
  static constexpr int STAT_OFF = P_OFF + 2 * SLAB_BYTES;    // float2 [2 item parities][2 column halves][128 rows]
  static constexpr int BAR_OFF = STAT_OFF + 2 * 2 * BM * 8;
  static constexpr int NUM_BARS = 20;
  static constexpr int TOTAL = BAR_OFF + NUM_BARS * 8 + 16 + 1024 /* alignment slack */;
  static constexpr int O_COL = S_BUFS * BN;                  // TMEM: S[0] | S[1] | S[2] | O
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// the two warps that share a row range (w and w + 4) meet on their own 64-thread named barrier: ids 2..5
__device__ __forceinline__ void pair_bar_sync(int wq) { asm volatile("bar.sync %0, 64;" ::"r"(2 + wq) : "memory"); }
// non-blocking probe of an mbarrier phase (the MMA thread serves two independent job streams)
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
      "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),
      "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

struct Item {
  int b, hh, qt, nblk;
  bool resident;            // scores (as exponentials) stay in TMEM between the passes: no second Q.K^T, no second ex2
  int qk_jobs;              // Q.K^T launches: nblk (resident) or 2 * nblk (recomputed in pass 2)
};
__device__ __forceinline__ Item decode_item(int idx, int bh_count, int H, int nq) {
  Item it;
  it.qt = nq - 1 - idx / bh_count;            // heaviest (most key blocks) first
  const int bh = idx - (idx / bh_count) * bh_count;
  it.b = bh / H;
  it.hh = bh - it.b * H;
  it.nblk = it.qt + 1;                        // causal: key blocks 0..qt
  it.resident = it.nblk <= RESIDENT_BLOCKS;
  it.qk_jobs = it.resident ? it.nblk : 2 * it.nblk;
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
  // barrier map (8 bytes each)
  auto q_full = [&](int s) { return bar_base + 8u * s; };
  auto q_empty = [&](int s) { return bar_base + 16 + 8u * s; };
  auto k_full = [&](int s) { return bar_base + 32 + 8u * s; };
  auto k_empty = [&](int s) { return bar_base + 48 + 8u * s; };
  auto s_full = [&](int s) { return bar_base + 64 + 8u * s; };      // 3
  auto s_empty = [&](int s) { return bar_base + 88 + 8u * s; };     // 3
  const uint32_t v_full = bar_base + 112, v_empty = bar_base + 120, p_full = bar_base + 128, p_empty = bar_base + 136,
                 o_full = bar_base + 144, o_empty = bar_base + 152;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + L::BAR_OFF + L::NUM_BARS * 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == QK_PRODUCER_WARP && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
    for (int s = 0; s < 2; ++s) {
      mbar_init(q_full(s), 1);
      mbar_init(q_empty(s), 1);
      mbar_init(k_full(s), 1);
      mbar_init(k_empty(s), 1);
    }
    for (int s = 0; s < S_BUFS; ++s) {
      mbar_init(s_full(s), 1);
      mbar_init(s_empty(s), 8);        // one arrival per softmax warp
    }
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
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

  auto load_tile = [&](const CUtensorMap* map, uint32_t dst, int col0, int row0, uint32_t bar) {
    mbar_expect_tx(bar, L::TILE_BYTES);
#pragma unroll
    for (int sl = 0; sl < L::SLABS; ++sl) tma_load_2d(dst + sl * L::SLAB_BYTES, map, col0 + sl * 64, row0, bar);
  };

  if (warp == QK_PRODUCER_WARP) {
    // ===================== TMA producer: Q (once per item, double-buffered) and the K tile of every Q.K^T job =====================
    if (lane == 0) {
      pdl_wait();                                   // q and the cache rows come from the QKV GEMM before us
      unsigned kcount = 0, n = 0;
      for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x, ++n) {
        const Item it = decode_item(idx, bh_count, H, nq);
        const int qb = n & 1;
        mbar_wait(q_empty(qb), ((n >> 1) & 1) ^ 1u);
        load_tile(&tmQ, smem_base + L::Q_OFF + qb * L::TILE_BYTES, it.hh * D, it.b * S + it.qt * BM, q_full(qb));
        for (int job = 0; job < it.qk_jobs; ++job, ++kcount) {
          const int j = job < it.nblk ? job : job - it.nblk;
          const int st = kcount & 1;
          mbar_wait(k_empty(st), ((kcount >> 1) & 1) ^ 1u);
          load_tile(&tmK, smem_base + L::K_OFF + st * L::TILE_BYTES, ((b0 + it.b) * H + it.hh) * D, j * BN, k_full(st));
        }
      }
    }
  } else if (warp == V_PRODUCER_WARP) {
    // ===================== TMA producer: the V tile of every P.V job (one stage: the next tile is not needed before the
    // softmax of its block is through) =====================
    if (lane == 0) {
      pdl_wait();
      unsigned vcount = 0;
      for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x) {
        const Item it = decode_item(idx, bh_count, H, nq);
        for (int j = 0; j < it.nblk; ++j, ++vcount) {
          mbar_wait(v_empty, (vcount & 1) ^ 1u);
          load_tile(&tmV, smem_base + L::V_OFF, ((b0 + it.b) * H + it.hh) * D, j * BN, v_full);
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================== MMA issuer: two independent job streams served by one thread =====================
    // Q.K^T jobs run ahead of the softmax as far as the three score buffers allow (also into the NEXT item); P.V jobs are
    // issued the moment the softmax warps have published a P tile.  Neither stream ever blocks the other: the thread
    // probes the barriers (mbarrier.test_wait) instead of waiting on them.
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc(BM, BN);                   // A = Q (K-major), B = K (K-major)
      constexpr uint32_t idesc_pv = make_idesc(BM, D) | (1u << 16);       // A = P (K-major), B = V (MN-major)
      int q_idx = blockIdx.x, p_idx = blockIdx.x;
      bool q_done = q_idx >= n_items, p_done = p_idx >= n_items;
      Item qit = decode_item(q_done ? 0 : q_idx, bh_count, H, nq), pit = qit;
      unsigned qn = 0, pn = 0, kc = 0, sc = 0, vc = 0, pc = 0;
      int q_job = 0, p_blk = 0;
      while (!q_done || !p_done) {
        bool progressed = false;
        if (!q_done) {
          const int ks = kc & 1, sb = sc % S_BUFS, qb = qn & 1;
          if ((q_job > 0 || mbar_test(q_full(qb), (qn >> 1) & 1)) && mbar_test(k_full(ks), (kc >> 1) & 1) &&
              mbar_test(s_empty(sb), ((sc / S_BUFS) & 1) ^ 1u)) {
            tcgen05_fence_after();
            const uint32_t qa = smem_base + L::Q_OFF + qb * L::TILE_BYTES, ka = smem_base + L::K_OFF + ks * L::TILE_BYTES;
#pragma unroll
            for (int kk = 0; kk < D / 16; ++kk) {
              const uint32_t off = (kk >> 2) * L::SLAB_BYTES + (kk & 3) * 32;   // 64-element slab; 16 elements = 32 bytes inside the swizzle row
              tcgen05_mma_f16(tmem_base + sb * BN, make_smem_desc(qa + off), make_smem_desc(ka + off), idesc_qk, kk > 0 ? 1u : 0u);
            }
            tcgen05_commit(k_empty(ks));
            tcgen05_commit(s_full(sb));
            ++kc;
            ++sc;
            progressed = true;
            if (++q_job == qit.qk_jobs) {
              tcgen05_commit(q_empty(qb));            // the Q buffer is free once this item's last Q.K^T has read it
              q_job = 0;
              ++qn;
              q_idx += gridDim.x;
              if (q_idx >= n_items) q_done = true;
              else qit = decode_item(q_idx, bh_count, H, nq);
            }
          }
        }
        if (!p_done) {
          if (mbar_test(p_full, pc & 1) && mbar_test(v_full, vc & 1) && (p_blk > 0 || mbar_test(o_empty, (pn & 1) ^ 1u))) {
            tcgen05_fence_after();
            const uint32_t pa = smem_base + L::P_OFF, va = smem_base + L::V_OFF;
#pragma unroll
            for (int kk = 0; kk < BN / 16; ++kk) {
              const uint64_t dp = make_smem_desc(pa + (kk >> 2) * L::SLAB_BYTES + (kk & 3) * 32);
              const uint64_t dv = make_smem_desc_mn(va + kk * 16 * 128, L::SLAB_BYTES);   // 16 keys = 16 rows of 128 bytes
              tcgen05_mma_f16(tmem_base + L::O_COL, dp, dv, idesc_pv, (p_blk > 0 || kk > 0) ? 1u : 0u);
            }
            tcgen05_commit(v_empty);
            tcgen05_commit(p_empty);
            ++vc;
            ++pc;
            progressed = true;
            if (++p_blk == pit.nblk) {
              tcgen05_commit(o_full);
              p_blk = 0;
              ++pn;
              p_idx += gridDim.x;
              if (p_idx >= n_items) p_done = true;
              else pit = decode_item(p_idx, bh_count, H, nq);
            }
          }
        }
        if (!progressed) __nanosleep(20);
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
    Item prev{};
    bool have_prev = false;

    // ctx = bf16(O) of the item whose P.V chain the MMA thread committed as number `on`
    auto epilogue = [&](const Item& it, unsigned on) {
      mbar_wait(o_full, on & 1);
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
    };
    // publish 64 probabilities (32 packed pairs) as this thread's row of the P tile's slab g
    auto publish_p = [&](const uint32_t* pk) {
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
    };

    for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x, ++n) {
      const Item it = decode_item(idx, bh_count, H, nq);
      const unsigned sc0 = sc;                      // first score buffer job of this item
      // ---------------- pass 1: m and l over this thread's 64 columns of every block
      float m = M_INIT, l = 0.f;
      float m_blk[RESIDENT_BLOCKS];                 // resident items: the running maximum each block's exponentials refer to
#pragma unroll
      for (int j = 0; j < RESIDENT_BLOCKS; ++j) m_blk[j] = M_INIT;
      for (int j = 0; j < it.nblk; ++j) {
        const int sb = sc % S_BUFS;
        mbar_wait(s_full(sb), (sc / S_BUFS) & 1);
        tcgen05_fence_after();
        uint32_t v[64];
        const uint32_t taddr = tmem_base + lane_addr + sb * BN + g * 64;
        tmem_ld32(taddr, v);
        tmem_ld32(taddr + 32, v + 32);
        tmem_ld_wait();
        if (!it.resident) {                         // scores are recomputed in pass 2: hand the buffer back now
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_empty(sb));
        }
        ++sc;
        float bmx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // four independent chains (latency, not issue, bounds this loop)
        if (j == it.qt) {                           // diagonal block: causal mask
#pragma unroll
          for (int i = 0; i < 64; i += 2) {
            float s0, s1;
            unpack_bf16x2(pack_bf16x2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), s0, s1);   // s = bf16(acc): the bmm's output
            if (g * 64 + i > row) s0 = -INFINITY;
            if (g * 64 + i + 1 > row) s1 = -INFINITY;
            v[i] = __float_as_uint(s0);
            v[i + 1] = __float_as_uint(s1);
            bmx[(i >> 1) & 3] = fmaxf(bmx[(i >> 1) & 3], fmaxf(s0, s1));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 64; i += 2) {
            float s0, s1;
            unpack_bf16x2(pack_bf16x2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), s0, s1);
            v[i] = __float_as_uint(s0);
            v[i + 1] = __float_as_uint(s1);
            bmx[(i >> 1) & 3] = fmaxf(bmx[(i >> 1) & 3], fmaxf(s0, s1));
          }
        }
        const float mn = fmaxf(fmaxf(m, fmaxf(bmx[0], bmx[1])), fmaxf(bmx[2], bmx[3]));
        float accx[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          const float e = ex2((__uint_as_float(v[i]) - mn) * LOG2E);
          accx[i & 7] += e;
          v[i] = __float_as_uint(e);
        }
        l = l * ex2((m - mn) * LOG2E) + (((accx[0] + accx[1]) + (accx[2] + accx[3])) + ((accx[4] + accx[5]) + (accx[6] + accx[7])));
        m = mn;
        if (it.resident) {                          // keep exp(s - m_j) where the scores were: pass 2 only rescales
          tmem_st32(taddr, v);
          tmem_st32(taddr + 32, v + 32);
#pragma unroll
          for (int jj = 0; jj < RESIDENT_BLOCKS; ++jj)
            if (jj == j) m_blk[jj] = mn;
        }
      }
      if (it.resident) tmem_st_wait();
      // the previous item's output leaves TMEM here, off the critical path (its P.V chain finished long ago)
      if (have_prev) epilogue(prev, n - 1);
      // the two column halves of a row meet: m = max, l rescaled to it
      stat[((n & 1) * 2 + g) * BM + row] = make_float2(m, l);
      pair_bar_sync(wq);
      const float2 o2 = stat[((n & 1) * 2 + (g ^ 1)) * BM + row];
      const float mf = fmaxf(m, o2.x);
      const float lf = l * ex2((m - mf) * LOG2E) + o2.y * ex2((o2.x - mf) * LOG2E);
      const float inv_l = 1.f / lf;
      // ---------------- pass 2: p = bf16(exp(s - m) / l) -> shared memory (this half's 64 keys = one K-major slab)
      if (it.resident) {
#pragma unroll
        for (int j = 0; j < RESIDENT_BLOCKS; ++j) {
          if (j < it.nblk) {
            const int sb = (sc0 + j) % S_BUFS;
            const uint32_t taddr = tmem_base + lane_addr + sb * BN + g * 64;
            uint32_t v[64];
            tmem_ld32(taddr, v);
            tmem_ld32(taddr + 32, v + 32);
            tmem_ld_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty(sb));
            const float f = ex2((m_blk[j] - mf) * LOG2E) * inv_l;
            uint32_t pk[32];
#pragma unroll
            for (int i = 0; i < 64; i += 2) pk[i >> 1] = pack_bf16x2(__uint_as_float(v[i]) * f, __uint_as_float(v[i + 1]) * f);
            publish_p(pk);
          }
        }
      } else {
        for (int j = 0; j < it.nblk; ++j) {
          const int sb = sc % S_BUFS;
          mbar_wait(s_full(sb), (sc / S_BUFS) & 1);
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
          publish_p(pk);
        }
      }
      prev = it;
      have_prev = true;
    }
    if (have_prev) epilogue(prev, n - 1);
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
