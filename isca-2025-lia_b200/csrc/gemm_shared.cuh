// Pieces of the tcgen05 GEMM shared by its two kernels: the single-GEMM kernel (gemm_sm100.cu) and the decode program
// kernel (decode_program_sm100.cu) -- tile constants, epilogue parameter blocks, the work scheduler (raster / stream-K),
// the rounding-point epilogue helpers and the SWAP-mode (decode) epilogue with its tensor-parallel exchange.
#pragma once
#include <stdlib.h>

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
  char* mc;   // multicast mapping of the arena (NVLS) or nullptr
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
    const int which = (n >= p.hq) + (n >= 2 * p.hq);   // n < 3 hq: no division (this runs once per row and 8 columns)
    const int c = n - which * p.hq;
    if (which == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = v[i] * p.q_scale;
      *reinterpret_cast<uint4*>(p.q_out + (size_t)m * p.hq + c) = pack8(v);
    } else {
      int bb = m, ss = 0;                              // decode appends one position per sequence
      if (p.S != 1) {
        bb = m / p.S;
        ss = m - bb * p.S;
      }
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
  if (p.mode == LIA_EPI_BIAS_RESIDUAL || p.mode == EPI_TP) return ldg_act(p.residual + (size_t)m * p.N + n);
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
  // `cta` / `ncta`: this CTA's index among the CTAs that share the GEMM and their number (default: the whole grid; the
  // decode program kernel runs GEMMs of fewer CTAs than its grid on a prefix of it)
  __device__ Sched(int k_blocks_, int tiles_a_, int tiles_b_, int streamk, bool pair = false, int cta_ = -1, int ncta_ = -1)
      : k_blocks(k_blocks_), tiles_a(tiles_a_), tiles_b(tiles_b_), step(pair ? gridDim.x >> 1 : gridDim.x), group_m(pair ? GROUP_M / 2 : GROUP_M) {
    const unsigned cta = cta_ < 0 ? blockIdx.x : (unsigned)cta_, ncta = ncta_ < 0 ? gridDim.x : (unsigned)ncta_;
    if (SWAP) {
      if (cta >= ncta) {
        pos = end = 0;
      } else if (streamk) {
        const unsigned total = (unsigned)tiles_a * (unsigned)k_blocks;
        pos = total * cta / ncta;
        end = total * (cta + 1) / ncta;
      } else {   // whole tiles only
        pos = ((unsigned)tiles_a * cta / ncta) * (unsigned)k_blocks;
        end = ((unsigned)tiles_a * (cta + 1) / ncta) * (unsigned)k_blocks;
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
// NVLink-switch multicast (NVLS): ONE address stands for the same location of every rank's arena.
//   multimem.ld_reduce  returns the SUM over all ranks' copies (8 bf16 per call, fp32 accumulation inside the switch)
//   multimem.st         stores to every rank's copy
__device__ __forceinline__ uint4 multimem_ld_reduce_bf16x8(const void* mc_addr) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(mc_addr)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_bf16x8(void* mc_addr, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.bf16x2 [%0], {%1,%2,%3,%4};" ::"l"(mc_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
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

// optional per-CTA timeline (LIA_GEMM_TRACE=1): 16 globaltimer stamps per CTA (<= 512 CTAs) in mapped host memory
__device__ __forceinline__ void stamp(unsigned long long* trace, int i) {
  if (trace != nullptr) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
    trace[blockIdx.x * 16 + i] = t;
  }
}

// Fused all-reduce, decode shapes, world >= 4: second half of the exchange (see the call site).  Kept out of line so
// that the plain projections -- which share this kernel image -- pay neither its registers nor its instruction bytes.
template <int BN>
__device__ __noinline__ void tp_two_shot_rows(const EpiParams& p, const uint4* rb, int epoch, int parity, int* tp_err, int ta,
                                              int m0, int n_col, int rows, unsigned long long* trace = nullptr) {
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
  unsigned mine = 0;                             // bit it: row group it of this tile is reduced here ((it + ta) % world == rank)
  {
    int o = ta % tp.world;
    for (int it = 0; it < ITERS; ++it) {
      if (o == tp.rank) mine |= 1u << it;
      o = o + 1 == tp.world ? 0 : o + 1;
    }
  }
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {           // rows this rank reduces
    const int m = m0 + it * 8;
    if (m >= rows) break;
    if (!((mine >> it) & 1u)) continue;
    const uint4 rsel = ldg_act(p.residual + (size_t)m * p.N + n_col);   // in flight while the partials are polled
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
  if (threadIdx.x == EPI_WARP0 * 32) stamp(trace, 10);
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
        if (m < rows && !((mine >> (it0 + j)) & 1u)) {
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
      if (m < rows && !((mine >> (it0 + j)) & 1u))
        *reinterpret_cast<uint4*>(p.out + (size_t)m * p.N + n_col) = make_uint4(lo[j].x, lo[j].z, hi[j].x, hi[j].z);
    }
  }
}



// ------------------------------------------------------------------ SWAP-mode (decode, M <= 128) epilogue of one work item
// Shared by the single-GEMM kernel (gemm_sm100.cu) and the decode program kernel (decode_program_sm100.cu): drains the
// accumulator (features x tokens), meets the other k-pieces of a split tile through the stream-K workspace in fixed k
// order, applies the epilogue and -- for the fused tensor-parallel projections -- runs the LL exchange.
struct SwapEpiCtx {
  float* stgf;                 // fp32 staging, BN x SWAP_LD floats
  float* ws;                   // stream-K pieces, one [BN][TILE_A] slot per CTA
  int* flags;                  // stream-K "piece published" flags, one per CTA
  unsigned long long* trace;
  int k_blocks, tiles_a;
  int cta, ncta;               // this CTA's index among the CTAs that share the GEMM, and their number
  int ew, lane, et;            // epilogue warp (TMEM lane quarter), lane, thread index among the 128 epilogue threads
  bool tp_on;
  int epoch, parity;
  int* tp_err;
};

template <int BN>
__device__ __forceinline__ void swap_epilogue_tile(const EpiParams& p, const SwapEpiCtx& cx, const Work& w, uint32_t taddr,
                                                   uint32_t tempty_addr, const float* biasf, const uint4* resv) {
  constexpr int ITERS = BN / 8;
  const TpDev& tp = p.tp;
  const int k_blocks = cx.k_blocks, tiles_a = cx.tiles_a;
  float* const ws = cx.ws;
  int* const flags = cx.flags;
  const int ew = cx.ew, lane = cx.lane, et = cx.et;
  const bool tp_on = cx.tp_on;
  const int epoch = cx.epoch, parity = cx.parity;
  int* const tp_err = cx.tp_err;
  const int c8 = et & 15, m0 = et >> 4;
  const int rows = min(BN, p.M);
  const int n_col = w.ta * TILE_A + c8 * 8;
  const bool n_ok = n_col < p.N;
  // tile rows = output features (TMEM lanes), columns = tokens: transpose through smem / workspace
  const int nl = ew * 32 + lane;                     // feature inside the tile
  float* stgf = cx.stgf;
  const bool full = (w.kb0 == 0 && w.kb1 == k_blocks);
  const bool owner = (w.kb0 == 0);                   // first k-piece: this CTA finishes the tile
  const bool two_shot = tp_on && (tp.opts & 128);    // fused all-reduce, world >= 4 (see the reduce side below)
  // two-shot: row group `it` (rows m0 + 8 it) of tile ta is reduced by rank (it + ta) % world.  Tracked incrementally as the
  // row groups are walked: a run-time modulo is ~30 dependent instructions, this code runs on one warp per scheduler with
  // nothing to hide them behind, and one modulo per row group and destination cost ~5 us per exchange (ncu source view)
  int own_rank = two_shot ? w.ta % tp.world : 0;
  float* slot = ws + (size_t)cx.cta * (BN * TILE_A);
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
  if (lane == 0) mbar_arrive(tempty_addr);
  if (!owner) {
    // publish this CTA's piece
    __threadfence();
    epi_bar_sync();
    if (et == 0) st_release_gpu(flags + cx.cta, 1);
  } else {
    // contributors are CTAs c+1, c+2, ... whose spans start inside this tile
    const unsigned total = (unsigned)tiles_a * (unsigned)k_blocks;
    const unsigned tile_end = (unsigned)(w.ta + 1) * (unsigned)k_blocks;
    int last_c = cx.cta;
    if (!full) {
      // the last CTA whose span starts inside this tile: start(c) = total * c / ncta < tile_end  <=>  c <= (tile_end * ncta - 1) / total
      // (one division instead of one per contributor: a tile of a small GEMM is split over up to ~8 CTAs)
      last_c = min(cx.ncta - 1, (int)((tile_end * (unsigned)cx.ncta - 1u) / total));
      last_c = max(last_c, cx.cta);
      // one thread per contributor polls its flag (acquire), the CTA barrier below hands what they acquired to everyone:
      // one thread polling them one after the other paid an L2 round trip per contributor even when all were long set
      for (int c = cx.cta + 1 + et; c <= last_c; c += 128)
        while (ld_acquire_gpu(flags + c) == 0) {
        }
    }
    epi_bar_sync();
    if (et == 0) stamp(cx.trace, 6);
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
        for (int c = cx.cta + 1; c <= last_c; ++c) {   // fixed k order: deterministic
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
              if (two_shot) {                                   // two-shot: only to the row group's reducing rank
                if (!(tp.opts & 8) || own_rank == tp.rank) {    // (bit 3: timing probe only, no remote stores)
                  uint4* dst = reinterpret_cast<uint4*>(tp.recv(own_rank, parity)) + idx;
                  st_volatile_v4(dst, lo);
                  st_volatile_v4(dst + 1, hi);
                }
              } else {
                for (int r2 = 0; r2 < tp.world; ++r2) {
                  if ((tp.opts & 8) && r2 != tp.rank) continue;
                  uint4* dst = reinterpret_cast<uint4*>(tp.recv(r2, parity)) + idx;
                  st_volatile_v4(dst, lo);
                  st_volatile_v4(dst + 1, hi);
                }
              }
            }
          }
          if (two_shot) own_rank = own_rank + 1 == tp.world ? 0 : own_rank + 1;
        }
        if (et == 0 && it0 == 0) stamp(cx.trace, 9);
      }
    }
    if (tp_on) {
      // reduce side: poll the `world` partials of this thread's values in OUR receive area until their
      // epoch words match, sum them in rank order (fp32, one rounding), add the residual: every rank
      // computes bit-identical results
      if (et == 0) stamp(cx.trace, 8);
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
          tp_two_shot_rows<BN>(p, rb, epoch, parity, tp_err, w.ta, m0, n_col, rows, cx.trace);
        }
      }
      if (et == 0) stamp(cx.trace, 11);
    }
    epi_bar_sync();                                  // staging reuse; all pieces consumed
    if (et == 0)
      for (int c = cx.cta + 1; c <= last_c; ++c) flags[c] = 0;   // re-arm for the next launch
  }
}

// ------------------------------------------------------------------ host side: tensor maps, launch plans, parameter blocks
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


// Argument validation + epilogue parameter block + launch plan of one projection: shared by lia_gemm_bf16 /
// lia_gemm_allreduce_bf16 and by the decode program builder, so that both produce IDENTICAL plans (same tiles, same
// stream-K spans, hence bit-identical results).
inline int gemm_fill_params(const void* A, const void* W, const void* bias, const void* residual, void* out, int M, int N, int K,
                            int epilogue, const LiaQkvArgs* qkv, const LiaTpArgs* tp, EpiParams& ep, Plan& pl) {

  const char* fn = tp ? "lia_gemm_allreduce_bf16" : "lia_gemm_bf16";
  LIA_CHECK_ARG(M > 0 && N > 0 && K > 0, "%s: M,N,K must be positive (got %d,%d,%d)", fn, M, N, K);
  LIA_CHECK_ARG(K % 8 == 0 && N % 8 == 0, "%s: K and N must be multiples of 8 (got K=%d N=%d)", fn, K, N);
  LIA_CHECK_ARG(A && W, "%s: null operand", fn);
  LIA_CHECK_ARG(((uintptr_t)A % 16 == 0) && ((uintptr_t)W % 16 == 0), "%s: operands must be 16-byte aligned", fn);
  ep = EpiParams{};
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
  pl = make_plan(M, N, K, /*allow_pair=*/tp == nullptr);
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
    ep.tp.mc = reinterpret_cast<char*>(tp->mc_arena);
    {
      // In-switch reduction moves S (every partial, once) + S/world bytes per rank against 2 S (world-1)/world for the
      // peer-store exchange: fewer from 4 ranks up (1.25 S vs 1.5 S; 1.125 S vs 1.75 S at 8), more at 2 (1.5 S vs S).
      // LIA_TP_NVLS=0 / 1 forces it off / on (A/B runs, and the 2-GPU test of the multimem path).
      const char* e = getenv("LIA_TP_NVLS");
      if (e ? atoi(e) == 0 : tp->world < 4) ep.tp.mc = nullptr;
    }
  }
  return LIA_OK;
}

}  // namespace
