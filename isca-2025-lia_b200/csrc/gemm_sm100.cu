// out[M,N] = epilogue(A[M,K] . W[N,K]^T), bf16 in / fp32 accumulate, for sm_100a.
//
// Replaces the reference's cuBLAS + ATen elementwise chains:
//   gpu_linear_compute*, gpu_linear_relu_compute*  (decoder.py:79-105)
//   the q/k/v projections, Q scaling and KV-cache writes (attentions.py:376-418,456-491)
//   the residual adds (decoder.py:229,310)
//
// Design (B200-first, not a translation of anything in the reference -- it has no GPU kernels):
//   * one persistent CTA per SM, 256 threads, warp-specialised:
//       warp 0  TMA producer   (cp.async.bulk.tensor, SWIZZLE_128B, mbarrier complete_tx)
//       warp 1  MMA issuer     (one elected lane issues tcgen05.mma.cta_group::1.kind::f16)
//       warp 2  TMEM allocator
//       warps 4-7 epilogue     (tcgen05.ld -> regs -> smem staging -> coalesced 16-byte global I/O)
//   * accumulators live in TMEM, double-buffered (2 x BN columns) so the epilogue of tile i
//     overlaps the main loop of tile i+1.
//   * both operands are K-major ([rows, K] row-major), so `x . W^T` needs no transpose:
//     W[N,K] is already the K-major "B" operand.
//   * two modes.  NORMAL (M > 128, prefill, tensor-bound): activations are the UMMA "A"
//     operand (128 rows per tile), W the "B" operand (BN = 128/256 rows).  SWAP (M <= 128,
//     decode, HBM-bound on W): W is the "A" operand so that its rows fill UMMA-M = 128, the
//     few tokens are UMMA-N (16..128); K can be split across CTAs so that every SM streams
//     weights; partial sums meet in an fp32 workspace and the last CTA to arrive reduces them
//     in a FIXED order (deterministic) and applies the epilogue.
//   * epilogue rounding points mirror the reference's eager ops exactly:
//       r1 = bf16(acc); r2 = bf16(r1 + bias); then relu | bf16(residual + r2) | bf16(r2*scale).
#include <cuda.h>

#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "common.cuh"
#include "umma.cuh"

namespace {

constexpr int BLOCK_K = 64;   // 64 bf16 = 128 bytes = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int TILE_A = 128;   // rows of the UMMA "A" operand per tile (UMMA M)
constexpr int NUM_THREADS = 256;
constexpr int EPI_WARP0 = 4;
constexpr int GROUP_M = 16;   // raster group (A tiles per group) for L2 reuse in NORMAL mode
constexpr int COUNTER_BYTES = 16384;
constexpr int SWAP_LD = TILE_A + 4;   // fp32 staging pitch (floats) in SWAP mode

constexpr int EPI_TP = 4;   // internal: row-parallel projection fused with its all-reduce + residual add
constexpr int TP_CTL_INTS = 64;                       // [0] epoch  [1] CTA exit counter  [2] error word
constexpr unsigned long long TP_TIMEOUT_NS = 4000000000ull;

struct TpDev {
  int rank, world;
  int opts;   // tuning probes: bit0 = do not trigger dependents early (PDL), bit1 = back off between failed polls
  char* arena[LIA_TP_MAX_WORLD];
  unsigned long long ctl_off, recv_off, recv_bytes, out_off;
  __device__ __forceinline__ int* ctl(int r) const { return reinterpret_cast<int*>(arena[r] + ctl_off); }
  // data_flag[unit][src]: rank `src`'s partial of `unit` has landed in rank r's receive area
  __device__ __forceinline__ int* data_flag(int r, int unit, int src) const {
    return ctl(r) + TP_CTL_INTS + unit * LIA_TP_MAX_WORLD + src;
  }
  // done_flag[unit]: the owner's final tile of `unit` has landed in rank r's `out`
  __device__ __forceinline__ int* done_flag(int r, int unit) const {
    return ctl(r) + TP_CTL_INTS + LIA_TP_MAX_UNITS * LIA_TP_MAX_WORLD + unit;
  }
  __device__ __forceinline__ bf16* recv(int r, int parity) const {
    return reinterpret_cast<bf16*>(arena[r] + recv_off + (unsigned long long)parity * recv_bytes);
  }
};

struct EpiParams {
  TpDev tp;
  const bf16* bias;
  const bf16* residual;
  bf16* out;
  int M, N;
  int mode;
  // LIA_EPI_QKV
  bf16* q_out;
  bf16* k_cache;
  bf16* v_cache;
  int hq, S, pos0, cache_batch, b0;
  float q_scale;
};

// ------------------------------------------------------------------ epilogue on 8 consecutive columns
// v[] holds r1 = bf16(acc) (as floats).  Rounding points follow SURVEY.md A.2.  The loads the
// epilogue needs (bias, residual) are split from the arithmetic so that callers can issue them
// early -- they do not depend on the accumulator.
__device__ __forceinline__ void epilogue_finish8(const EpiParams& p, int m, int n, float* v, const float* b,
                                                 const uint4& res) {
  if (p.bias != nullptr) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = bf16r(v[i] + b[i]);
  }
  if (p.mode == LIA_EPI_BIAS) {
    *reinterpret_cast<uint4*>(p.out + (size_t)m * p.N + n) = pack8(v);
  } else if (p.mode == LIA_EPI_BIAS_RELU) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
    *reinterpret_cast<uint4*>(p.out + (size_t)m * p.N + n) = pack8(v);
  } else if (p.mode == LIA_EPI_BIAS_RESIDUAL || p.mode == EPI_TP) {
    float r[8];
    unpack8(res, r);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = r[i] + v[i];
    *reinterpret_cast<uint4*>(p.out + (size_t)m * p.N + n) = pack8(v);
  } else {  // LIA_EPI_QKV
    const int which = n / p.hq;
    const int c = n - which * p.hq;
    if (which == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = v[i] * p.q_scale;
      *reinterpret_cast<uint4*>(p.q_out + (size_t)m * p.hq + c) = pack8(v);
    } else {
      const int bb = m / p.S;
      const int ss = m - bb * p.S;
      bf16* cache = (which == 1) ? p.k_cache : p.v_cache;
      const size_t row = (size_t)(p.pos0 + ss) * p.cache_batch + p.b0 + bb;
      *reinterpret_cast<uint4*>(cache + row * p.hq + c) = pack8(v);
    }
  }
}
__device__ __forceinline__ void epilogue_load_bias8(const EpiParams& p, int n, float* b) {
  if (p.bias != nullptr) unpack8(__ldg(reinterpret_cast<const uint4*>(p.bias + n)), b);
}
__device__ __forceinline__ uint4 epilogue_load_residual8(const EpiParams& p, int m, int n) {
  if (p.mode == LIA_EPI_BIAS_RESIDUAL || p.mode == EPI_TP) return ldg_stream(p.residual + (size_t)m * p.N + n);
  return make_uint4(0, 0, 0, 0);
}
__device__ __forceinline__ void epilogue_store8(const EpiParams& p, int m, int n, float* v) {
  float b[8];
  epilogue_load_bias8(p, n, b);
  const uint4 res = epilogue_load_residual8(p, m, n);
  epilogue_finish8(p, m, n, v, b, res);
}

// PAIR: two CTAs (one cluster, one TPC) compute a 256 x BN tile with tcgen05.mma.cta_group::2; each CTA stages its own
// 128 rows of A and only HALF of the B tile, so a stage is 32 KB instead of 48 KB and six stages fit.
template <bool SWAP, int BN, int STAGES, bool PAIR = false>
struct SmemLayout {
  static constexpr int A_BYTES = TILE_A * BLOCK_K * 2;
  static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGING_BYTES = SWAP ? BN * SWAP_LD * 4 : 4 * 32 * 128;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES + STAGING_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16 + 1024 /* alignment slack */;
};

__host__ __device__ constexpr int tmem_cols(int bn) { return 2 * bn <= 32 ? 32 : 2 * bn <= 64 ? 64 : 2 * bn <= 128 ? 128 : 2 * bn <= 256 ? 256 : 512; }

// ------------------------------------------------------------------ work scheduling
// NORMAL: unit u = blockIdx.x + i*gridDim.x over (A tile, B tile) pairs, rastered in groups of
//         GROUP_M A-tiles so that concurrently running CTAs share operands through L2.
// SWAP  : "stream-K".  The iteration space is the flat list of (W row-tile, k-block) pairs; CTA c
//         owns the contiguous span [c*total/G, (c+1)*total/G), so EVERY SM streams the same number
//         of weight bytes whatever N and K are.  A span covers at most one trailing piece of a
//         tile (kb0 > 0: written to this CTA's fp32 workspace slot), whole tiles, and at most one
//         leading piece (kb0 == 0, kb1 < k_blocks: this CTA owns the tile and adds the pieces of
//         CTAs c+1, c+2, ... in k order -- a fixed order, so results are deterministic).
struct Work {
  int ta, tb, kb0, kb1;
};

template <bool SWAP>
struct Sched {
  // 32-bit arithmetic throughout: tiles_a * k_blocks * gridDim.x < 2^31 is checked on the host (64-bit
  // divisions are software routines of ~100 instructions each and this code is inlined into three roles)
  int k_blocks, tiles_a, tiles_b;
  unsigned pos, end;   // SWAP: flat k-block position; NORMAL: unit index / count
  unsigned step;       // NORMAL: workers (CTAs, or CTA pairs) sharing the unit list
  int group_m;         // NORMAL: A tiles per raster group
  __device__ Sched(int k_blocks_, int tiles_a_, int tiles_b_, int streamk, bool pair = false)
      : k_blocks(k_blocks_), tiles_a(tiles_a_), tiles_b(tiles_b_), step(pair ? gridDim.x >> 1 : gridDim.x), group_m(pair ? GROUP_M / 2 : GROUP_M) {
    if (SWAP) {
      if (streamk) {
        const unsigned total = (unsigned)tiles_a * (unsigned)k_blocks;
        pos = total * blockIdx.x / gridDim.x;
        end = total * (blockIdx.x + 1) / gridDim.x;
      } else {   // whole tiles only
        pos = ((unsigned)tiles_a * blockIdx.x / gridDim.x) * (unsigned)k_blocks;
        end = ((unsigned)tiles_a * (blockIdx.x + 1) / gridDim.x) * (unsigned)k_blocks;
      }
    } else {
      pos = pair ? blockIdx.x >> 1 : blockIdx.x;
      end = (unsigned)tiles_a * (unsigned)tiles_b;
    }
  }
  __device__ bool next(Work& w) {
    if (pos >= end) return false;
    if (SWAP) {
      w.ta = (int)(pos / (unsigned)k_blocks);
      w.tb = 0;
      w.kb0 = (int)(pos - (unsigned)w.ta * (unsigned)k_blocks);
      const unsigned left = end - pos;
      w.kb1 = (left < (unsigned)(k_blocks - w.kb0)) ? w.kb0 + (int)left : k_blocks;
      pos += (unsigned)(w.kb1 - w.kb0);
    } else {
      const int u = (int)pos;
      const int group_size = group_m * tiles_b;
      const int group = u / group_size;
      const int first = group * group_m;
      const int gm = min(group_m, tiles_a - first);
      const int r = u - group * group_size;
      w.ta = first + r % gm;
      w.tb = r / gm;
      w.kb0 = 0;
      w.kb1 = k_blocks;
      pos += step;
    }
    return true;
  }
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_volatile_v4(uint4* p, const uint4& v) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_volatile_v4(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
// wait until *flag reaches `epoch` (flags only grow; wrap-safe compare).  A peer that never shows up
// must not hang the GPU: after TP_TIMEOUT_NS the error word is set and every later wait falls through.
__device__ __noinline__ void tp_spin(const int* flag, int epoch, int* err) {
  unsigned long long t0 = 0;
  unsigned it = 0;
  while ((int)(ld_acquire_sys(flag) - epoch) < 0) {
    if ((++it & 255u) == 0) {
      if (*reinterpret_cast<volatile int*>(err) != 0) return;
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
      if (t0 == 0) t0 = t;
      else if (t - t0 > TP_TIMEOUT_NS) {
        atomicExch(err, 1);
        return;
      }
    }
  }
}

// Fused all-reduce, decode shapes, world >= 4: second half of the exchange (see the call site).  Kept out of line so
// that the plain projections -- which share this kernel image -- pay neither its registers nor its instruction bytes.
template <int BN>
__device__ __noinline__ void tp_two_shot_rows(const EpiParams& p, const uint4* rb, int epoch, int parity, int* tp_err, int ta,
                                              int m0, int n_col, int rows) {
  constexpr int ITERS = BN / 8;
  const TpDev& tp = p.tp;
  // ---- two-shot (world >= 4).  Ownership is per (row group, tile): the 8 values of row m = m0 + 8*it of
  // tile ta are reduced by rank (it + ta) % world, so EVERY thread of EVERY CTA on every rank reduces 1/world
  // of its own values (all `world` partials in flight at once: one L2 round trip) and receives the rest as
  // finals (all in flight at once: one more round trip) -- two one-way NVLink latencies per exchange,
  // 2(world-1)/world x the data per rank, and no CTA idles while a few "owner" CTAs reduce whole tiles.
  const size_t final_base = ((size_t)tp.world * BN * p.N) >> 2;   // finals follow the `world` partial slots
  unsigned long long t0 = 0;
  unsigned spins = 0;
  auto give_up = [&]() -> bool {                 // a peer that never shows up must not hang the GPU
    if (tp.opts & 4) return true;                // timing probe only: do not wait (results are garbage)
    if ((++spins & 255u) != 0) return false;
    if (*reinterpret_cast<volatile int*>(tp_err) != 0) return true;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
    if (t0 == 0) t0 = t;
    else if (t - t0 > TP_TIMEOUT_NS) {
      atomicExch(tp_err, 1);
      return true;
    }
    return false;
  };
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {           // rows this rank reduces
    const int m = m0 + it * 8;
    if (m >= rows) break;
    if ((it + ta) % tp.world != tp.rank) continue;
    const uint4 rsel = ldg_stream(p.residual + (size_t)m * p.N + n_col);   // in flight while the partials are polled
    const uint4* q0 = rb + (((size_t)m * p.N + n_col) >> 2);
    const size_t src_stride = ((size_t)BN * p.N) >> 2;
    float sum[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) sum[i] = 0.f;
    constexpr int SH = LIA_TP_MAX_WORLD / 2;     // partials in flight per poll (keeps the caller spill-free)
#pragma unroll 1
    for (int s0 = 0; s0 < tp.world; s0 += SH) {
      uint4 lo[SH], hi[SH];
      bool ok;
      do {
        ok = true;
#pragma unroll
        for (int k = 0; k < SH; ++k)
          if (s0 + k < tp.world) {
            lo[k] = ld_volatile_v4(q0 + (s0 + k) * src_stride);
            hi[k] = ld_volatile_v4(q0 + (s0 + k) * src_stride + 1);
          }
#pragma unroll
        for (int k = 0; k < SH; ++k)
          if (s0 + k < tp.world)
            ok = ok && lo[k].y == (uint32_t)epoch && lo[k].w == (uint32_t)epoch && hi[k].y == (uint32_t)epoch &&
                 hi[k].w == (uint32_t)epoch;
      } while (!ok && !give_up());
#pragma unroll
      for (int k = 0; k < SH; ++k)
        if (s0 + k < tp.world) {                 // rank order, fp32: bit-identical on every rank
          float g[8];
          unpack_bf16x2(lo[k].x, g[0], g[1]);
          unpack_bf16x2(lo[k].z, g[2], g[3]);
          unpack_bf16x2(hi[k].x, g[4], g[5]);
          unpack_bf16x2(hi[k].z, g[6], g[7]);
#pragma unroll
          for (int i = 0; i < 8; ++i) sum[i] += g[i];
        }
    }
    float r[8];
    unpack8(rsel, r);
#pragma unroll
    for (int i = 0; i < 8; ++i) sum[i] = r[i] + bf16r(sum[i]);
    const uint4 o = pack8(sum);
    const uint4 flo = make_uint4(o.x, (uint32_t)epoch, o.y, (uint32_t)epoch);
    const uint4 fhi = make_uint4(o.z, (uint32_t)epoch, o.w, (uint32_t)epoch);
    const size_t fidx = final_base + (((size_t)m * p.N + n_col) >> 2);
    for (int r2 = 0; r2 < tp.world; ++r2) {      // second shot: finals to every peer, framed like the partials
      if (r2 == tp.rank || (tp.opts & 8)) continue;
      uint4* dst = reinterpret_cast<uint4*>(tp.recv(r2, parity)) + fidx;
      st_volatile_v4(dst, flo);
      st_volatile_v4(dst + 1, fhi);
    }
    *reinterpret_cast<uint4*>(p.out + (size_t)m * p.N + n_col) = o;
  }
  constexpr int RG = ITERS < 4 ? ITERS : 4;      // rows other ranks reduce: their finals, RG rows in flight
#pragma unroll 1
  for (int it0 = 0; it0 < ITERS; it0 += RG) {
    uint4 lo[RG], hi[RG];
    bool ok;
    do {
      ok = true;
#pragma unroll
      for (int j = 0; j < RG; ++j) {
        const int m = m0 + (it0 + j) * 8;
        if (m < rows && (it0 + j + ta) % tp.world != tp.rank) {
          const uint4* q = rb + final_base + (((size_t)m * p.N + n_col) >> 2);
          lo[j] = ld_volatile_v4(q);
          hi[j] = ld_volatile_v4(q + 1);
          ok = ok && lo[j].y == (uint32_t)epoch && lo[j].w == (uint32_t)epoch && hi[j].y == (uint32_t)epoch &&
               hi[j].w == (uint32_t)epoch;
        }
      }
    } while (!ok && !give_up());
#pragma unroll
    for (int j = 0; j < RG; ++j) {
      const int m = m0 + (it0 + j) * 8;
      if (m < rows && (it0 + j + ta) % tp.world != tp.rank)
        *reinterpret_cast<uint4*>(p.out + (size_t)m * p.N + n_col) = make_uint4(lo[j].x, lo[j].z, hi[j].x, hi[j].z);
    }
  }
}

// optional per-CTA timeline (LIA_GEMM_TRACE=1): 16 globaltimer stamps per CTA (<= 512 CTAs) in mapped host memory
__device__ __forceinline__ void stamp(unsigned long long* trace, int i) {
  if (trace != nullptr) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
    trace[blockIdx.x * 16 + i] = t;
  }
}

// ------------------------------------------------------------------ the kernel
template <bool SWAP, int BN, int STAGES, bool TP, bool PAIR = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
lia_gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const __grid_constant__ EpiParams p,
                        int k_blocks, int streamk, int tiles_a, int tiles_b, float* __restrict__ ws,
                        int* __restrict__ flags, unsigned long long* __restrict__ trace) {
  static_assert(!PAIR || (!SWAP && !TP && (BN == 256 || BN == 224)), "CTA pairs: plain prefill projections with 256/224-wide tiles only");
  using L = SmemLayout<SWAP, BN, STAGES, PAIR>;
  // TMEM columns from one accumulator to the other: BN, rounded up to the epilogue's 64-column step for tile widths that
  // are not a multiple of it (BN = 224, the opt-in tile that trims the wave tail of the N = 7168 projections)
  constexpr int ACC_COLS = (SWAP || BN % 64 == 0) ? BN : ((BN + 63) / 64) * 64;
  // PAIR: rank in the 2-CTA cluster; rank 0 (the "leader") issues every MMA and owns the full / tempty barriers
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + L::BAR_OFFSET;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + L::BAR_OFFSET + (2 * STAGES + 4) * 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) stamp(trace, 0);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);      // PAIR: only the leader's is used; it collects the bytes of both CTAs' loads
      mbar_init(empty_bar(s), 1);     // PAIR: the leader's commit arrives on both CTAs' copies
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), PAIR ? 8 : 4);   // PAIR: the epilogue warps of both CTAs release the leader's accumulator
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if (PAIR) {   // collective over the pair: one warp of EACH CTA issues it
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr_smem)),
                   "r"((uint32_t)tmem_cols(BN))
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr_smem)),
                   "r"((uint32_t)tmem_cols(BN))
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // the peer's barriers are initialised before anything arrives on them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (threadIdx.x == 0) stamp(trace, 1);
  // PDL: from here on the next kernel in the stream may become resident; this kernel touches nothing a
  // predecessor produced until pdl_wait() (only weight tiles are fetched before it, see the producer)
  if (!((SWAP ? p.mode == EPI_TP : TP) && (p.tp.opts & 1))) pdl_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool first_issue = true;
      int npre = 0;
      if (SWAP) {
        // operand A is the weight matrix, which no kernel writes: fill the pipeline with weight tiles
        // BEFORE waiting for the previous kernel (whose output is operand B)
        Sched<SWAP> pre(k_blocks, tiles_a, tiles_b, streamk);
        Work w;
        while (npre < STAGES && pre.next(w)) {
          for (int kb = w.kb0; kb < w.kb1 && npre < STAGES; ++kb, ++npre) {
            mbar_expect_tx(full_bar(npre), L::STAGE_BYTES);
            tma_load_2d(smem_base + npre * L::STAGE_BYTES, &tmA, kb * BLOCK_K, w.ta * TILE_A, full_bar(npre));
          }
        }
      }
      pdl_wait();
      Sched<SWAP> sched(k_blocks, tiles_a, tiles_b, streamk, PAIR);
      Work w;
      int idx = 0;
      while (sched.next(w)) {
        for (int kb = w.kb0; kb < w.kb1; ++kb, ++idx) {
          const uint32_t sa = smem_base + stage * L::STAGE_BYTES;
          const uint32_t sb = sa + L::A_BYTES;
          if (PAIR) {
            // this CTA's 128 rows of the 256-row A tile and its half of the B tile; every byte is credited to the
            // LEADER's full barrier, which expects both CTAs' stages (the leader alone arrives on it)
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t fb = mapa_u32(full_bar(stage), 0);
            if (cta_rank == 0) mbar_expect_tx(full_bar(stage), 2 * L::STAGE_BYTES);
            tma_load_2d_pair(sa, &tmA, kb * BLOCK_K, (w.ta * 2 + (int)cta_rank) * TILE_A, fb);
            tma_load_2d_pair(sb, &tmB, kb * BLOCK_K, w.tb * BN + (int)cta_rank * (BN / 2), fb);
          } else if (idx < npre) {
            tma_load_2d(sb, &tmB, kb * BLOCK_K, w.tb * BN, full_bar(stage));   // A tile already in flight
          } else {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            mbar_expect_tx(full_bar(stage), L::STAGE_BYTES);
            tma_load_2d(sa, &tmA, kb * BLOCK_K, w.ta * TILE_A, full_bar(stage));
            tma_load_2d(sb, &tmB, kb * BLOCK_K, w.tb * BN, full_bar(stage));
          }
          if (first_issue) {
            stamp(trace, 2);
            first_issue = false;
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = make_idesc(PAIR ? 2 * TILE_A : TILE_A, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      bool first_full = true;
      Sched<SWAP> sched(k_blocks, tiles_a, tiles_b, streamk, PAIR);
      Work w;
      while (sched.next(w)) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_COLS);
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          if (first_full) {
            stamp(trace, 3);
            first_full = false;
          }
          tcgen05_fence_after();
          const uint32_t sa = smem_base + stage * L::STAGE_BYTES;
          const uint64_t da = make_smem_desc(sa);
          const uint64_t db = make_smem_desc(sa + L::A_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 16 elements = 32 bytes along K inside the 128-byte swizzle row: +2 in the >>4 address field
            if (PAIR) tcgen05_mma_f16_pair(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > w.kb0 || k > 0) ? 1u : 0u);
            else tcgen05_mma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > w.kb0 || k > 0) ? 1u : 0u);
          }
          if (PAIR) tcgen05_commit_pair(empty_bar(stage));   // frees the stage in BOTH CTAs
          else tcgen05_commit(empty_bar(stage));   // frees the smem stage when these MMAs have read it
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (PAIR) tcgen05_commit_pair(tfull_bar(acc));   // each CTA's epilogue drains its own 128 rows
        else tcgen05_commit(tfull_bar(acc));       // accumulator complete -> epilogue
        stamp(trace, 4);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue =====================
    const int ew = warp - EPI_WARP0;          // == warp % 4 -> TMEM lanes [32*ew, 32*ew+32)
    const int et = threadIdx.x - EPI_WARP0 * 32;
    pdl_wait();                               // residual / outputs / stream-K workspace belong to the stream order
    int acc = 0;
    uint32_t acc_phase = 0;
    // tensor-parallel fused all-reduce: this launch's epoch (same on every rank: all ranks issue the same
    // call sequence) selects the receive-area parity and is the value flags are raised to
    // NORMAL (prefill): TP is a separate instantiation, the plain projections carry none of this code.
    // SWAP (decode): ONE function serves plain and fused projections -- a decode step alternates between them
    // every few microseconds, and two ~100 KB kernels evicting each other from the instruction cache cost
    // ~10 us per switch (measured); a run-time branch keeps the whole step on one resident code image.
    const bool tp_on = SWAP ? (p.mode == EPI_TP) : TP;
    const TpDev& tp = p.tp;
    int epoch = 0, parity = 0;
    int* tp_err = nullptr;
    if (tp_on) {
      epoch = (tp.opts & 64) ? 1 : *reinterpret_cast<volatile int*>(tp.ctl(tp.rank)) + 1;
      parity = epoch & 1;
      tp_err = tp.ctl(tp.rank) + 2;
    }
    int pend_u = -1, pend_ta = 0, pend_tb = 0;   // NORMAL two-shot: one owned tile whose reduction is deferred
    // NORMAL two-shot: flags are raised one tile late.  A flag may only be raised after a system-scope fence has
    // seen the remote stores it covers acknowledged (~5 us over NVLink); a tile later the acknowledgements are
    // long back, so the fence costs nothing and the epilogue warps never sit out an NVLink round trip.
    int sig_data_u = -1, sig_done_u = -1;
    auto tp_flush_signals = [&]() {
      if (sig_data_u >= 0 || sig_done_u >= 0) {
        __threadfence_system();
        epi_bar_sync();
        if (sig_data_u >= 0 && et == 0) st_release_sys(tp.data_flag(sig_data_u % tp.world, sig_data_u, tp.rank), epoch);
        if (sig_done_u >= 0 && et < tp.world && et != tp.rank) st_release_sys(tp.done_flag(et, sig_done_u), epoch);
        sig_data_u = sig_done_u = -1;
      }
    };
    // NORMAL two-shot, owner side: all partials of tile u are here -> reduce in rank order, add the
    // residual, write the final tile into EVERY rank's `out`, then raise done_flag[u] on the peers
    auto tp_reduce_owned = [&](int u, int ta, int tb) {
      if (et < tp.world && et != tp.rank && !(tp.opts & 4)) tp_spin(tp.data_flag(tp.rank, u, et), epoch, tp_err);
      epi_bar_sync();
      const bf16* rbase = tp.recv(tp.rank, parity) + (size_t)(u / tp.world) * tp.world * (TILE_A * BN);
      constexpr int CPR = BN / 8;             // 16-byte chunks per tile row
      constexpr int CH = 8;                   // chunks per thread in flight: (1 + world) x 8 independent 16-byte loads
      static_assert(SWAP || !TP || (TILE_A * CPR) % (128 * CH) == 0, "tile must split into whole batches");
#pragma unroll 1
      for (int c0 = et; c0 < TILE_A * CPR; c0 += 128 * CH) {
        float sum[CH][8];
        uint4 res[CH];
        bool ok[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          const int c = c0 + j * 128;
          const int row = c / CPR, ch = c - row * CPR;
          const int m = ta * TILE_A + row, n = tb * BN + ch * 8;
          ok[j] = (m < p.M && n < p.N);
          res[j] = ok[j] ? ldg_stream(p.residual + (size_t)m * p.N + n) : make_uint4(0, 0, 0, 0);
#pragma unroll
          for (int i = 0; i < 8; ++i) sum[j][i] = 0.f;
        }
        for (int src = 0; src < tp.world; ++src) {
          uint4 v[CH];
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const int c = c0 + j * 128;
            const int row = c / CPR, ch = c - row * CPR;
            v[j] = ok[j] ? __ldcg(reinterpret_cast<const uint4*>(rbase + ((size_t)src * TILE_A + row) * BN + ch * 8)) : make_uint4(0, 0, 0, 0);
          }
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            float f[8];
            unpack8(v[j], f);
#pragma unroll
            for (int i = 0; i < 8; ++i) sum[j][i] += f[i];
          }
        }
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          if (ok[j]) {
            const int c = c0 + j * 128;
            const int row = c / CPR, ch = c - row * CPR;
            const int m = ta * TILE_A + row, n = tb * BN + ch * 8;
            float r[8];
            unpack8(res[j], r);
#pragma unroll
            for (int i = 0; i < 8; ++i) sum[j][i] = r[i] + bf16r(sum[j][i]);
            const uint4 o = pack8(sum[j]);
            const size_t off = tp.out_off + ((size_t)m * p.N + n) * 2;
            for (int r2 = 0; r2 < tp.world; ++r2) *reinterpret_cast<uint4*>(tp.arena[r2] + off) = o;
          }
        }
      }
      sig_done_u = u;                         // done_flag[u] is raised by the next tp_flush_signals()
    };
    Sched<SWAP> sched(k_blocks, tiles_a, tiles_b, streamk, PAIR);
    const int row_tile = PAIR ? 2 : 1;        // 128-row blocks per A tile; this CTA drains block `cta_rank`
    Work w;
    while (sched.next(w)) {
      // SWAP: this thread finishes columns [n_col, n_col+8) of rows m0, m0+8, ... -- fetch what the
      // epilogue needs besides the accumulator (bias, residual) BEFORE waiting for the MMAs
      if (tp_on && !SWAP) tp_flush_signals();
      constexpr int ITERS = SWAP ? BN / 8 : 1;
      const int c8 = et & 15, m0 = et >> 4;
      const int rows = min(BN, p.M);
      const int n_col = w.ta * TILE_A + c8 * 8;
      const bool n_ok = SWAP && n_col < p.N;
      float biasf[8];
      uint4 resv[ITERS];
      if (SWAP && w.kb0 == 0 && n_ok) {
        epilogue_load_bias8(p, n_col, biasf);
        const bool need_res = !(tp_on && (tp.opts & 128));   // two-shot: only the reducing rank of a row reads its residual
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
          const int m = m0 + it * 8;
          resv[it] = (m < rows && need_res) ? epilogue_load_residual8(p, m, n_col) : make_uint4(0, 0, 0, 0);
        }
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      if (et == 0) stamp(trace, 5);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(acc * ACC_COLS) + ((uint32_t)(ew * 32) << 16);

      if (!SWAP) {
        // tile rows = tokens (TMEM lanes), columns = output features
        const uint32_t stg = smem_base + STAGES * L::STAGE_BYTES + ew * 4096;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 64) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            if (BN % 64 != 0 && c0 + half * 32 >= BN) continue;   // (compile-time false for 128/256-wide tiles)
            uint32_t v[32];
            tmem_ld32(taddr + c0 + half * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int chunk = half * 4 + j;
              const uint32_t addr = stg + lane * 128 + ((chunk ^ (lane & 7)) << 4);
              const uint32_t x0 = pack_bf16x2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
              const uint32_t x1 = pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
              const uint32_t x2 = pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
              const uint32_t x3 = pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(x0), "r"(x1), "r"(x2), "r"(x3) : "memory");
            }
          }
          if (c0 + 64 >= BN) {
            // all TMEM reads of this accumulator are done: hand it back to the MMA warp early
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (PAIR && cta_rank != 0) mbar_arrive_cluster(mapa_u32(tempty_bar(acc), 0));
              else mbar_arrive(tempty_bar(acc));
            }
          } else {
            __syncwarp();
          }
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + (lane >> 3);
            const int ch = lane & 7;
            uint4 q;
            const uint32_t addr = stg + r * 128 + ((ch ^ (r & 7)) << 4);
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(addr));
            const int m = (w.ta * row_tile + (int)cta_rank) * TILE_A + ew * 32 + r;
            const int n = w.tb * BN + c0 + ch * 8;
            if (m < p.M && n < p.N && (BN % 64 == 0 || c0 + ch * 8 < BN)) {
              float f[8];
              unpack8(q, f);
              if (tp_on) {
                // partial tile (r2 = bf16(bf16(acc) + bias/world)) -> the owner's receive area, slot [tile][this rank]
                if (p.bias != nullptr) {
                  float b[8];
                  epilogue_load_bias8(p, n, b);
#pragma unroll
                  for (int i = 0; i < 8; ++i) f[i] = bf16r(f[i] + b[i]);
                }
                const int u = w.ta * tiles_b + w.tb;
                bf16* dst = tp.recv(u % tp.world, parity) + ((size_t)(u / tp.world) * tp.world + tp.rank) * (TILE_A * BN) +
                            (size_t)(ew * 32 + r) * BN + c0 + ch * 8;
                *reinterpret_cast<uint4*>(dst) = pack8(f);
              } else {
                epilogue_store8(p, m, n, f);
              }
            }
          }
          __syncwarp();
        }
        if (tp_on) {
          const int u = w.ta * tiles_b + w.tb;
          const int owner = u % tp.world;
          if (owner != tp.rank) {
            sig_data_u = u;
          } else {
            epi_bar_sync();                      // our own partial is complete in our receive area
            // defer the reduction by one owned tile (= `world` tiles of MMA work) so the peers' partials are
            // normally already here and the epilogue warps never stall the tensor pipe on NVLink latency
            if (pend_u >= 0) tp_reduce_owned(pend_u, pend_ta, pend_tb);
            pend_u = u; pend_ta = w.ta; pend_tb = w.tb;
          }
        }
      } else {
        // tile rows = output features (TMEM lanes), columns = tokens: transpose through smem / workspace
        const int nl = ew * 32 + lane;                     // feature inside the tile
        float* stgf = reinterpret_cast<float*>(smem_gen + STAGES * L::STAGE_BYTES);
        const bool full = (w.kb0 == 0 && w.kb1 == k_blocks);
        const bool owner = (w.kb0 == 0);                   // first k-piece: this CTA finishes the tile
        const bool two_shot = tp_on && (tp.opts & 128);    // fused all-reduce, world >= 4 (see the reduce side below)
        float* slot = ws + (size_t)blockIdx.x * (BN * TILE_A);
        constexpr int CH = BN >= 32 ? 32 : 16;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += CH) {
          uint32_t v[CH];
          if (CH == 32) tmem_ld32(taddr + c0, v); else tmem_ld16(taddr + c0, v);
          tmem_ld_wait();
          if (owner) {
#pragma unroll
            for (int j = 0; j < CH; ++j) stgf[(c0 + j) * SWAP_LD + nl] = __uint_as_float(v[j]);
          } else {
#pragma unroll
            for (int j = 0; j < CH; ++j)
              if (c0 + j < p.M) slot[(c0 + j) * TILE_A + nl] = __uint_as_float(v[j]);
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
        if (!owner) {
          // publish this CTA's piece
          __threadfence();
          epi_bar_sync();
          if (et == 0) st_release_gpu(flags + blockIdx.x, 1);
        } else {
          // contributors are CTAs c+1, c+2, ... whose spans start inside this tile
          const unsigned total = (unsigned)tiles_a * (unsigned)k_blocks;
          const unsigned tile_end = (unsigned)(w.ta + 1) * (unsigned)k_blocks;
          int last_c = blockIdx.x;
          if (!full) {
            while (last_c + 1 < (int)gridDim.x && total * (unsigned)(last_c + 1) / gridDim.x < tile_end) ++last_c;
            if (et == 0) {
              for (int c = blockIdx.x + 1; c <= last_c; ++c)
                while (ld_acquire_gpu(flags + c) == 0) {
                }
            }
          }
          epi_bar_sync();
          if (et == 0) stamp(trace, 6);
          // thread -> 8 fixed columns (c8) and rows m0, m0+8, ...: ROWS_PER_BATCH rows at a time so that
          // the loads of a batch (own sums from smem, pieces from L2) are all in flight together
          if (n_ok) {
            constexpr int RB = ITERS < 4 ? ITERS : 4;
            // (a real loop, not unrolled: this code runs once per tile and its size is paid in instruction-cache
            // misses -- the unrolled epilogue was ~120 KB of SASS)
#pragma unroll 1
            for (int it0 = 0; it0 < ITERS; it0 += RB) {
              uint4 rcur[RB];      // this batch's prefetched residual rows (register select, no dynamic indexing)
#pragma unroll
              for (int j = 0; j < RB; ++j) {
                rcur[j] = resv[j];
#pragma unroll
                for (int b = 1; b < ITERS / RB; ++b)
                  if (it0 == b * RB) rcur[j] = resv[b * RB + j];
              }
              float f[RB][8];
#pragma unroll
              for (int j = 0; j < RB; ++j) {
                const int m = m0 + (it0 + j) * 8;
                const float4 a = *reinterpret_cast<const float4*>(stgf + m * SWAP_LD + c8 * 8);
                const float4 b = *reinterpret_cast<const float4*>(stgf + m * SWAP_LD + c8 * 8 + 4);
                f[j][0] = a.x; f[j][1] = a.y; f[j][2] = a.z; f[j][3] = a.w;
                f[j][4] = b.x; f[j][5] = b.y; f[j][6] = b.z; f[j][7] = b.w;
              }
              for (int c = blockIdx.x + 1; c <= last_c; ++c) {   // fixed k order: deterministic
                float4 pa[RB], pb[RB];
#pragma unroll
                for (int j = 0; j < RB; ++j) {
                  const int m = m0 + (it0 + j) * 8;
                  const float* src = ws + (size_t)c * (BN * TILE_A) + m * TILE_A + c8 * 8;
                  if (m < rows) {
                    pa[j] = __ldcg(reinterpret_cast<const float4*>(src));
                    pb[j] = __ldcg(reinterpret_cast<const float4*>(src + 4));
                  } else {
                    pa[j] = pb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                  }
                }
#pragma unroll
                for (int j = 0; j < RB; ++j) {
                  f[j][0] += pa[j].x; f[j][1] += pa[j].y; f[j][2] += pa[j].z; f[j][3] += pa[j].w;
                  f[j][4] += pb[j].x; f[j][5] += pb[j].y; f[j][6] += pb[j].z; f[j][7] += pb[j].w;
                }
              }
#pragma unroll
              for (int j = 0; j < RB; ++j) {
                const int m = m0 + (it0 + j) * 8;
                if (m < rows) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) f[j][i] = bf16r(f[j][i]);
                  if (!tp_on || (tp.opts & 32)) {
                    epilogue_finish8(p, m, n_col, f[j], biasf, rcur[j]);
                  } else {
                    // one-shot all-reduce, push side: this rank's partial (r2 = bf16(bf16(acc) + bias/world)) goes
                    // into slot [this rank] of EVERY rank's receive area (peers over NVLink, fire-and-forget).
                    // "LL" framing: every 8-byte word is {4 bytes of data, epoch}, so data and its validity
                    // arrive together -- no fence, no separate flag, ONE one-way NVLink latency per exchange.
                    if (p.bias != nullptr) {
#pragma unroll
                      for (int i = 0; i < 8; ++i) f[j][i] = bf16r(f[j][i] + biasf[i]);
                    }
                    const uint4 o = pack8(f[j]);
                    const uint4 lo = make_uint4(o.x, (uint32_t)epoch, o.y, (uint32_t)epoch);
                    const uint4 hi = make_uint4(o.z, (uint32_t)epoch, o.w, (uint32_t)epoch);
                    const size_t idx = (((size_t)tp.rank * BN + m) * p.N + n_col) >> 2;   // uint4 index: 2 per 8 values
                    for (int r2 = 0; r2 < tp.world; ++r2) {
                      if ((tp.opts & 8) && r2 != tp.rank) continue;   // timing probe only: no remote stores
                      if (two_shot && r2 != (it0 + j + w.ta) % tp.world) continue;   // two-shot: only the row's reducing rank
                      uint4* dst = reinterpret_cast<uint4*>(tp.recv(r2, parity)) + idx;
                      st_volatile_v4(dst, lo);
                      st_volatile_v4(dst + 1, hi);
                    }
                  }
                }
              }
            }
          }
          if (tp_on) {
            // reduce side: poll the `world` partials of this thread's values in OUR receive area until their
            // epoch words match, sum them in rank order (fp32, one rounding), add the residual: every rank
            // computes bit-identical results
            if (et == 0) stamp(trace, 8);
            if (n_ok && !(tp.opts & 32)) {
              const uint4* rb = reinterpret_cast<const uint4*>(tp.recv(tp.rank, parity));
              if (!two_shot) {
                constexpr int RB2 = ITERS < 4 ? ITERS : 4;
  #pragma unroll 1
                for (int it0 = 0; it0 < ITERS; it0 += RB2) {
                  uint4 rcur[RB2];
  #pragma unroll
                  for (int j = 0; j < RB2; ++j) {
                    rcur[j] = resv[j];
  #pragma unroll
                    for (int b = 1; b < ITERS / RB2; ++b)
                      if (it0 == b * RB2) rcur[j] = resv[b * RB2 + j];
                  }
                  float sum[RB2][8];
  #pragma unroll
                  for (int j = 0; j < RB2; ++j)
  #pragma unroll
                    for (int i = 0; i < 8; ++i) sum[j][i] = 0.f;
                  for (int src = 0; src < tp.world; ++src) {
                    uint4 lo[RB2], hi[RB2];
                    unsigned long long t0 = 0;
                    unsigned spins = 0;
                    bool ok;
                    do {
                      ok = true;
  #pragma unroll
                      for (int j = 0; j < RB2; ++j) {
                        const int m = m0 + (it0 + j) * 8;
                        if (m < rows) {
                          const uint4* q = rb + ((((size_t)src * BN + m) * p.N + n_col) >> 2);
                          lo[j] = ld_volatile_v4(q);
                          hi[j] = ld_volatile_v4(q + 1);
                        }
                      }
  #pragma unroll
                      for (int j = 0; j < RB2; ++j) {
                        const int m = m0 + (it0 + j) * 8;
                        if (m < rows)
                          ok = ok && lo[j].y == (uint32_t)epoch && lo[j].w == (uint32_t)epoch && hi[j].y == (uint32_t)epoch &&
                               hi[j].w == (uint32_t)epoch;
                      }
                      if (tp.opts & 4) ok = true;               // timing probe only: do not wait for the peer (results are garbage)
                      if (!ok && (tp.opts & 2)) __nanosleep(100);
                      if (!ok && (++spins & 255u) == 0) {       // a peer that never shows up must not hang the GPU
                        if (*reinterpret_cast<volatile int*>(tp_err) != 0) break;
                        unsigned long long t;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
                        if (t0 == 0) t0 = t;
                        else if (t - t0 > TP_TIMEOUT_NS) {
                          atomicExch(tp_err, 1);
                          break;
                        }
                      }
                    } while (!ok);
  #pragma unroll
                    for (int j = 0; j < RB2; ++j) {
                      float g[8];
                      unpack_bf16x2(lo[j].x, g[0], g[1]);
                      unpack_bf16x2(lo[j].z, g[2], g[3]);
                      unpack_bf16x2(hi[j].x, g[4], g[5]);
                      unpack_bf16x2(hi[j].z, g[6], g[7]);
  #pragma unroll
                      for (int i = 0; i < 8; ++i) sum[j][i] += g[i];
                    }
                  }
  #pragma unroll
                  for (int j = 0; j < RB2; ++j) {
                    const int m = m0 + (it0 + j) * 8;
                    if (m < rows) {
                      float r[8];
                      unpack8(rcur[j], r);
  #pragma unroll
                      for (int i = 0; i < 8; ++i) sum[j][i] = r[i] + bf16r(sum[j][i]);
                      *reinterpret_cast<uint4*>(p.out + (size_t)m * p.N + n_col) = pack8(sum[j]);
                    }
                  }
                }
              } else {
                tp_two_shot_rows<BN>(p, rb, epoch, parity, tp_err, w.ta, m0, n_col, rows);
              }
            }
            if (et == 0) stamp(trace, 11);
          }
          epi_bar_sync();                                  // staging reuse; all pieces consumed
          if (et == 0)
            for (int c = blockIdx.x + 1; c <= last_c; ++c) flags[c] = 0;   // re-arm for the next launch
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (tp_on) {
      if (!SWAP) {
        tp_flush_signals();
        if (pend_u >= 0) tp_reduce_owned(pend_u, pend_ta, pend_tb);
        tp_flush_signals();
        // this CTA's tiles that other ranks own: wait until their final values have landed in our `out`
        // (the kernel must not complete before its output is complete); 128 threads poll in parallel
        Sched<SWAP> s2(k_blocks, tiles_a, tiles_b, streamk);
        Work w2;
        int i = 0;
        while (s2.next(w2)) {
          const int u = w2.ta * tiles_b + w2.tb;
          if (u % tp.world != tp.rank && (i++ & 127) == et && !(tp.opts & 4)) tp_spin(tp.done_flag(tp.rank, u), epoch, tp_err);
        }
        epi_bar_sync();
      }
      // the last CTA to leave publishes the epoch for the next launch (every CTA read it on entry)
      if (et == 0 && !(tp.opts & 16)) {
        __threadfence();
        int* ctl = tp.ctl(tp.rank);
        if (atomicAdd(ctl + 1, 1) == (int)gridDim.x - 1) {
          ctl[1] = 0;
          __threadfence();
          *reinterpret_cast<volatile int*>(ctl) = epoch;
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // no CTA leaves (or frees TMEM) while its peer may still arrive on its barriers
  if (threadIdx.x == 0) stamp(trace, 7);
  if (warp == 2) {
    tcgen05_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)tmem_cols(BN)) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)tmem_cols(BN)) : "memory");
  }
}

// ------------------------------------------------------------------ host side
// 2-D bf16 tensor [rows, K] row-major, box = [box_rows, 64] with 128-byte swizzle; OOB -> zeros
int make_tmap(CUtensorMap* map, const void* base, int rows, int K, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    lia_set_error("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    return LIA_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    lia_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%d K=%d box_rows=%d)", (int)r, rows, K, box_rows);
    return LIA_ERR_CUDA;
  }
  return LIA_OK;
}

struct Plan {
  bool swap;
  bool pair;     // NORMAL: 2-CTA clusters, tcgen05.mma.cta_group::2 on 256 x 256 tiles
  int bn;
  int grid;      // CTAs
  int streamk;   // SWAP: split tiles across CTAs at k-block granularity
  int tiles_a, tiles_b, k_blocks;
};

// CTAs available to one GEMM launch: the SM count, or LIA_GEMM_MAX_CTAS when set (lets two launches that
// talk to each other share one GPU: tests/test_gpu_tp.py runs a 2-rank exchange on a single device)
int gemm_cta_budget() {
  int sms = lia_sm_count();
  const char* env = getenv("LIA_GEMM_MAX_CTAS");
  if (env) {
    const int v = atoi(env);
    if (v > 0 && v < sms) sms = v;
  }
  return sms;
}

// The CTA-pair prefill kernel is the default (measured on a B200: bit-identical to the one-CTA kernel and 6-12 % faster on
// the OPT-30B prefill shapes, profiles/README.md); LIA_GEMM_2CTA=0 selects the one-CTA 128 x 256 kernel for A/B runs.
bool pair_enabled() {
  const char* env = getenv("LIA_GEMM_2CTA");
  return env == nullptr || atoi(env) != 0;
}

bool bn224_enabled() {
  const char* env = getenv("LIA_GEMM_BN224");
  return env != nullptr && atoi(env) != 0;
}

Plan make_plan(int M, int N, int K, bool allow_pair = false) {
  Plan pl;
  pl.pair = false;
  const int sms = gemm_cta_budget();
  pl.k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
  pl.swap = (M <= 128);
  pl.streamk = 0;
  if (pl.swap) {
    pl.bn = M <= 16 ? 16 : M <= 32 ? 32 : M <= 64 ? 64 : 128;
    pl.tiles_a = (N + TILE_A - 1) / TILE_A;
    pl.tiles_b = 1;
    // stream-K: every CTA streams the same number of k-blocks (>= 4 each)
    const long long total = (long long)pl.tiles_a * pl.k_blocks;
    long long g = total / 4;
    if (g < 1) g = 1;
    if (g > sms) g = sms;
    pl.grid = (int)g;
    pl.streamk = (total % pl.grid != 0 || pl.tiles_a % pl.grid != 0) ? 1 : 0;
    const char* env = getenv("LIA_STREAMK");   // tuning/debug override: 0 = whole tiles per CTA
    if (env && atoi(env) == 0) pl.streamk = 0;
    if (!pl.streamk && pl.grid > pl.tiles_a) pl.grid = pl.tiles_a;
  } else {
    pl.bn = (N % 256 == 0 || N >= 2048) ? 256 : 128;
    pl.pair = allow_pair && pl.bn == 256 && M >= 4 * TILE_A && sms >= 2 && pair_enabled();
    pl.tiles_a = pl.pair ? (M + 2 * TILE_A - 1) / (2 * TILE_A) : (M + TILE_A - 1) / TILE_A;
    if (pl.pair && N % 224 == 0 && bn224_enabled()) {
      // opt-in (LIA_GEMM_BN224=1, until A/B-checked on hardware): 256 x 224 tiles where they shorten the schedule.
      // N = 7168, M = 8192 on 74 CTA pairs: 896 tiles of 256 = 12.1 waves -> 13 x 256 columns of work per pair;
      // 1024 tiles of 224 = 13.8 waves -> 14 x 224 = 5.8 % less.  Same K order per element: bit-identical results.
      const long long pairs = sms / 2;
      const long long u256 = (long long)pl.tiles_a * ((N + 255) / 256), u224 = (long long)pl.tiles_a * (N / 224);
      if (((u224 + pairs - 1) / pairs) * 224 < ((u256 + pairs - 1) / pairs) * 256) pl.bn = 224;
    }
    pl.tiles_b = (N + pl.bn - 1) / pl.bn;
    const int units = pl.tiles_a * pl.tiles_b;
    if (pl.pair) pl.grid = 2 * (units < sms / 2 ? units : sms / 2);
    else pl.grid = units < sms ? units : sms;
  }
  return pl;
}

size_t plan_workspace(const Plan& pl) {
  if (!(pl.swap && pl.streamk)) return COUNTER_BYTES;
  return COUNTER_BYTES + (size_t)pl.grid * pl.bn * TILE_A * sizeof(float);
}

static unsigned g_trace_seq = 0;
// debug timeline buffer (mapped pinned host memory): 64 launches x 1024 CTAs x 8 stamps, only when LIA_GEMM_TRACE is set
// The stamps live in DEVICE memory (stamping into mapped host memory made every traced kernel wait out
// PCIe write acknowledgements at its end); lia_debug_gemm_trace() copies them to a host mirror.
constexpr size_t TRACE_WORDS = (size_t)64 * 512 * 16;
static unsigned long long* g_trace_host = nullptr;
unsigned long long* trace_buffer() {
  static unsigned long long* buf = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* env = getenv("LIA_GEMM_TRACE");
    if (env && atoi(env) != 0) {
      void* p = nullptr;
      if (cudaMalloc(&p, TRACE_WORDS * sizeof(unsigned long long)) == cudaSuccess) {
        cudaMemset(p, 0, TRACE_WORDS * sizeof(unsigned long long));
        buf = reinterpret_cast<unsigned long long*>(p);
        g_trace_host = reinterpret_cast<unsigned long long*>(calloc(TRACE_WORDS, sizeof(unsigned long long)));
      }
    }
  }
  return buf;
}

// CTA-pair launch: clusters of two CTAs (one TPC) + programmatic dependent launch
template <int BN>
int launch_pair(const Plan& pl, const CUtensorMap& tmA, const CUtensorMap& tmB, const EpiParams& ep, cudaStream_t stream) {
  constexpr int STAGES = 6;
  using L = SmemLayout<false, BN, STAGES, true>;
  static_assert(L::TOTAL <= 232448, "shared memory budget exceeded");
  auto kern = lia_gemm_tcgen05_kernel<false, BN, STAGES, false, true>;
  static bool configured = false;
  if (!configured) {
    LIA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(pl.grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = L::TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // persistent schedule: never launch more clusters than can be co-resident (a TPC with one SM fused off holds none)
  static int max_clusters = 0;
  if (max_clusters == 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = lia_sm_count() / 2;
    }
    max_clusters = n;
  }
  if (pl.grid > 2 * max_clusters) cfg.gridDim = dim3(2 * max_clusters);
  cfg.numAttrs = lia_pdl_enabled() ? 2 : 1;
  float* ws = nullptr;
  int* flags = nullptr;
  unsigned long long* tr = nullptr;
  LIA_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, ep, pl.k_blocks, pl.streamk, pl.tiles_a, pl.tiles_b, ws, flags, tr));
  return LIA_OK;
}

template <bool SWAP, int BN, int STAGES, bool TP>
int launch_tp(const Plan& pl, const CUtensorMap& tmA, const CUtensorMap& tmB, const EpiParams& ep, float* ws, int* flags,
           cudaStream_t stream) {
  using L = SmemLayout<SWAP, BN, STAGES>;
  static_assert(L::TOTAL <= 232448, "shared memory budget exceeded");
  auto kern = lia_gemm_tcgen05_kernel<SWAP, BN, STAGES, TP>;
  static bool configured = false;
  if (!configured) {
    LIA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  unsigned long long* tr = trace_buffer() ? trace_buffer() + (size_t)(g_trace_seq++ % 64) * 512 * 16 : nullptr;
  LIA_CUDA(lia_launch(kern, dim3(pl.grid), dim3(NUM_THREADS), L::TOTAL, stream, tmA, tmB, ep, pl.k_blocks, pl.streamk,
                      pl.tiles_a, pl.tiles_b, ws, flags, tr));
  return LIA_OK;
}

template <bool SWAP, int BN, int STAGES>
int launch(const Plan& pl, const CUtensorMap& tmA, const CUtensorMap& tmB, const EpiParams& ep, float* ws, int* flags,
           cudaStream_t stream) {
  if (!SWAP && ep.mode == EPI_TP) return launch_tp<SWAP, BN, STAGES, true>(pl, tmA, tmB, ep, ws, flags, stream);
  return launch_tp<SWAP, BN, STAGES, false>(pl, tmA, tmB, ep, ws, flags, stream);
}

}  // namespace

__global__ void debug_marker_kernel(unsigned long long* dst) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  *dst = t;
}
// debug only: a 1-thread kernel that writes %globaltimer into slot `idx` of the marker area (last 1024 entries)
extern "C" int lia_debug_marker(int idx, void* stream) {
  unsigned long long* buf = trace_buffer();
  if (!buf) return -1;
  debug_marker_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(buf + 63 * 512 * 16 + idx);
  return 0;
}

// debug only (not part of include/lia_b200.h): host pointer to the last launch's timeline
extern "C" const unsigned long long* lia_debug_gemm_trace(void) {
  if (!trace_buffer() || !g_trace_host) return nullptr;
  cudaDeviceSynchronize();
  cudaMemcpy(g_trace_host, trace_buffer(), TRACE_WORDS * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  return g_trace_host;
}

extern "C" size_t lia_gemm_workspace_bytes(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  return plan_workspace(make_plan(M, N, K));
}

static int gemm_impl(const void* A, const void* W, const void* bias, const void* residual, void* out, int M, int N, int K,
                     int epilogue, const LiaQkvArgs* qkv, const LiaTpArgs* tp, void* workspace, size_t workspace_bytes,
                     cudaStream_t stream) {
  const char* fn = tp ? "lia_gemm_allreduce_bf16" : "lia_gemm_bf16";
  LIA_CHECK_ARG(M > 0 && N > 0 && K > 0, "%s: M,N,K must be positive (got %d,%d,%d)", fn, M, N, K);
  LIA_CHECK_ARG(K % 8 == 0 && N % 8 == 0, "%s: K and N must be multiples of 8 (got K=%d N=%d)", fn, K, N);
  LIA_CHECK_ARG(A && W, "%s: null operand", fn);
  LIA_CHECK_ARG(((uintptr_t)A % 16 == 0) && ((uintptr_t)W % 16 == 0), "%s: operands must be 16-byte aligned", fn);
  EpiParams ep{};
  ep.bias = reinterpret_cast<const bf16*>(bias);
  ep.residual = reinterpret_cast<const bf16*>(residual);
  ep.out = reinterpret_cast<bf16*>(out);
  ep.M = M;
  ep.N = N;
  ep.mode = epilogue;
  if (epilogue == LIA_EPI_QKV) {
    LIA_CHECK_ARG(qkv != nullptr, "lia_gemm_bf16: LIA_EPI_QKV needs LiaQkvArgs");
    LIA_CHECK_ARG(qkv->hq > 0 && qkv->hq % 8 == 0 && N == 3 * qkv->hq, "lia_gemm_bf16: QKV needs N == 3*hq, hq %% 8 == 0");
    LIA_CHECK_ARG(qkv->S > 0 && M % qkv->S == 0, "lia_gemm_bf16: QKV needs M %% S == 0");
    LIA_CHECK_ARG(qkv->q_out && qkv->k_cache && qkv->v_cache, "lia_gemm_bf16: QKV null output");
    LIA_CHECK_ARG(qkv->b0 >= 0 && qkv->b0 + M / qkv->S <= qkv->cache_batch && qkv->pos0 >= 0, "lia_gemm_bf16: QKV batch window");
    ep.q_out = reinterpret_cast<bf16*>(qkv->q_out);
    ep.k_cache = reinterpret_cast<bf16*>(qkv->k_cache);
    ep.v_cache = reinterpret_cast<bf16*>(qkv->v_cache);
    ep.hq = qkv->hq; ep.S = qkv->S; ep.pos0 = qkv->pos0; ep.cache_batch = qkv->cache_batch; ep.b0 = qkv->b0;
    ep.q_scale = qkv->q_scale;
  } else {
    LIA_CHECK_ARG(out != nullptr, "%s: null output", fn);
    if (epilogue == LIA_EPI_BIAS_RESIDUAL || epilogue == EPI_TP)
      LIA_CHECK_ARG(residual != nullptr, "%s: residual epilogue needs residual", fn);
  }
  Plan pl = make_plan(M, N, K, /*allow_pair=*/tp == nullptr);
  LIA_CHECK_ARG((long long)pl.tiles_a * pl.k_blocks * (pl.grid + 1) < (1ll << 31) && (long long)pl.tiles_a * pl.tiles_b < (1ll << 31),
                "%s: problem too large for the 32-bit tile scheduler (M=%d N=%d K=%d)", fn, M, N, K);
  if (tp != nullptr) {
    LIA_CHECK_ARG(tp->world >= 2 && tp->world <= LIA_TP_MAX_WORLD && tp->rank >= 0 && tp->rank < tp->world,
                  "%s: bad rank/world %d/%d", fn, tp->rank, tp->world);
    for (int r = 0; r < tp->world; ++r) LIA_CHECK_ARG(tp->arena[r] != nullptr, "%s: arena[%d] is not mapped", fn, r);
    LIA_CHECK_ARG(tp->ctl_off % 16 == 0 && tp->recv_off % 16 == 0 && tp->recv_bytes % 16 == 0, "%s: arena offsets must be 16-byte aligned", fn);
    LIA_CHECK_ARG(tp->recv_bytes >= lia_tp_recv_bytes(M, N, K, tp->world), "%s: receive area of %llu bytes is too small (need %zu)", fn,
                  (unsigned long long)tp->recv_bytes, lia_tp_recv_bytes(M, N, K, tp->world));
    const int units = pl.swap ? pl.tiles_a : pl.tiles_a * pl.tiles_b;
    LIA_CHECK_ARG(units <= LIA_TP_MAX_UNITS, "%s: %d output tiles exceed LIA_TP_MAX_UNITS", fn, units);
    if (!pl.swap) {
      const char* base = reinterpret_cast<const char*>(tp->arena[tp->rank]);
      LIA_CHECK_ARG(reinterpret_cast<const char*>(out) == base + tp->out_off, "%s: for M > 128 `out` must live in the arena at out_off", fn);
    }
    {
      const char* e1 = getenv("LIA_TP_LATE_TRIGGER");
      const char* e2 = getenv("LIA_TP_POLL_BACKOFF");
      const char* e3 = getenv("LIA_TP_NO_WAIT");
      const char* e4 = getenv("LIA_TP_NO_PUSH");
      const char* e5 = getenv("LIA_TP_OPTS");        // raw probe bits (16: skip the exit accounting)
      if (e5) ep.tp.opts |= atoi(e5);
      // decode (M <= 128): one-shot costs (world-1) x the data per rank and one NVLink hop, two-shot 2(world-1)/world x
      // and two hops -- measured cross-over between world 2 and 4
      const char* e6 = getenv("LIA_TP_DECODE_TWOSHOT");
      if (pl.swap && (e6 ? atoi(e6) != 0 : tp->world >= 4)) ep.tp.opts |= 128;
      ep.tp.opts |= ((e1 && atoi(e1)) ? 1 : 0) | ((e2 && atoi(e2)) ? 2 : 0) | ((e3 && atoi(e3)) ? 4 : 0) | ((e4 && atoi(e4)) ? 8 : 0);
    }
    ep.tp.rank = tp->rank;
    ep.tp.world = tp->world;
    for (int r = 0; r < tp->world; ++r) ep.tp.arena[r] = reinterpret_cast<char*>(tp->arena[r]);
    ep.tp.ctl_off = tp->ctl_off; ep.tp.recv_off = tp->recv_off; ep.tp.recv_bytes = tp->recv_bytes; ep.tp.out_off = tp->out_off;
  }
  float* ws = nullptr;
  int* flags = nullptr;
  if (pl.swap && pl.streamk) {
    if (workspace == nullptr || workspace_bytes < plan_workspace(pl) || pl.grid * (int)sizeof(int) > COUNTER_BYTES) {
      pl.streamk = 0;   // no workspace: whole tiles per CTA (some SMs idle) rather than fail
      if (pl.grid > pl.tiles_a) pl.grid = pl.tiles_a;
    } else {
      flags = reinterpret_cast<int*>(workspace);
      ws = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + COUNTER_BYTES);
    }
  }
  CUtensorMap tmA, tmB;
  int rc;
  if (pl.swap) {
    if ((rc = make_tmap(&tmA, W, N, K, TILE_A)) != LIA_OK) return rc;
    if ((rc = make_tmap(&tmB, A, M, K, pl.bn)) != LIA_OK) return rc;
    switch (pl.bn) {
      case 16: return launch<true, 16, 8>(pl, tmA, tmB, ep, ws, flags, stream);
      case 32: return launch<true, 32, 8>(pl, tmA, tmB, ep, ws, flags, stream);
      case 64: return launch<true, 64, 8>(pl, tmA, tmB, ep, ws, flags, stream);
      default: return launch<true, 128, 4>(pl, tmA, tmB, ep, ws, flags, stream);
    }
  } else {
    if ((rc = make_tmap(&tmA, A, M, K, TILE_A)) != LIA_OK) return rc;
    if ((rc = make_tmap(&tmB, W, N, K, pl.pair ? pl.bn / 2 : pl.bn)) != LIA_OK) return rc;
    if (pl.pair) return pl.bn == 224 ? launch_pair<224>(pl, tmA, tmB, ep, stream) : launch_pair<256>(pl, tmA, tmB, ep, stream);
    if (pl.bn == 256) return launch<false, 256, 4>(pl, tmA, tmB, ep, ws, flags, stream);
    return launch<false, 128, 6>(pl, tmA, tmB, ep, ws, flags, stream);
  }
}

extern "C" int lia_gemm_bf16(const void* A, const void* W, const void* bias, const void* residual, void* out, int M,
                             int N, int K, int epilogue, const LiaQkvArgs* qkv, void* workspace, size_t workspace_bytes,
                             lia_stream_t stream_) {
  LIA_CHECK_ARG(epilogue >= LIA_EPI_BIAS && epilogue <= LIA_EPI_QKV, "lia_gemm_bf16: unknown epilogue %d", epilogue);
  return gemm_impl(A, W, bias, residual, out, M, N, K, epilogue, qkv, nullptr, workspace, workspace_bytes,
                   reinterpret_cast<cudaStream_t>(stream_));
}

// ------------------------------------------------------------------ tensor-parallel fused projection + all-reduce
extern "C" size_t lia_tp_ctl_bytes(void) {
  return (size_t)(TP_CTL_INTS + LIA_TP_MAX_UNITS * LIA_TP_MAX_WORLD + LIA_TP_MAX_UNITS) * sizeof(int);
}

extern "C" size_t lia_tp_recv_bytes(int M, int N, int K, int world) {
  if (M <= 0 || N <= 0 || K <= 0 || world <= 0) return 0;
  const Plan pl = make_plan(M, N, K);
  if (pl.swap) return (size_t)(world + 1) * pl.bn * N * sizeof(bf16) * 2;             // [src][bn rows][N] partials + [bn][N] finals, {data, epoch} words
  const size_t units = (size_t)pl.tiles_a * pl.tiles_b;
  return ((units + world - 1) / world) * world * (size_t)(TILE_A * pl.bn) * sizeof(bf16);   // [owned tile][src][128][bn]
}

extern "C" int lia_gemm_allreduce_bf16(const void* A, const void* W, const void* bias, const void* residual, void* out,
                                       int M, int N, int K, const LiaTpArgs* tp, void* workspace, size_t workspace_bytes,
                                       lia_stream_t stream_) {
  LIA_CHECK_ARG(tp != nullptr, "lia_gemm_allreduce_bf16: null LiaTpArgs");
  return gemm_impl(A, W, bias, residual, out, M, N, K, EPI_TP, nullptr, tp, workspace, workspace_bytes,
                   reinterpret_cast<cudaStream_t>(stream_));
}

extern "C" int lia_tp_error(const LiaTpArgs* tp) {
  LIA_CHECK_ARG(tp != nullptr && tp->rank >= 0 && tp->rank < LIA_TP_MAX_WORLD && tp->arena[tp->rank] != nullptr, "lia_tp_error: bad LiaTpArgs");
  LIA_CUDA(cudaDeviceSynchronize());
  int* err = reinterpret_cast<int*>(reinterpret_cast<char*>(tp->arena[tp->rank]) + tp->ctl_off) + 2;
  int v = 0;
  LIA_CUDA(cudaMemcpy(&v, err, sizeof(int), cudaMemcpyDeviceToHost));
  if (v != 0) {
    LIA_CUDA(cudaMemset(err, 0, sizeof(int)));
    lia_set_error("tensor-parallel exchange timed out waiting for a peer (error word %d)", v);
    return 1;
  }
  return 0;
}
