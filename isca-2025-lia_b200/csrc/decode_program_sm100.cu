// One decode step (or any chain of decode-shaped operations, M <= 128 rows) as ONE persistent kernel.
//
// A decode step of the reference is ~30 library launches per layer; the first-generation path here was
// 7 launches per layer (LayerNorm, QKV GEMM, attention, out_proj, LayerNorm, fc1, fc2).  Each of those kernels streams its
// weights at ~98 % of the HBM peak INSIDE its main loop but pays ~12 us of fixed cost around it (launch, prologue, first
// tile latency, stream-K hand-off, drain) during which HBM idles -- 0.75 of the roofline for the step.  Here the whole
// chain is a "program" of operations executed by one grid of persistent CTAs (one per SM):
//
//   * the TMA producer warp never stops: while an operation's epilogue and the grid-wide dependency on it are still being
//     resolved, it is already streaming the NEXT operation's weight tiles (which depend on nothing) into the shared-memory
//     ring, and -- for attention -- the cached K/V rows of earlier positions (immutable during the step);
//   * operations are ordered by ONE monotonic arrival counter in global memory: a CTA that has finished its part of
//     operation i adds 1; whoever needs the outputs of operations < i waits until the counter reaches i * gridDim;
//   * GEMMs are exactly the stream-K swap-AB tcgen05 GEMM of gemm_sm100.cu (same tiles, same spans, same fixed-order
//     reduction: gemm_shared.cuh), LayerNorm / embedding / argmax are the bodies of the stand-alone kernels
//     (small_ops.cuh), and decode attention repeats attn_decode.cu's arithmetic in the same order on K/V tiles that arrive
//     through the TMA ring instead of through per-thread loads -- so a program's results are BIT-IDENTICAL to the
//     kernel-per-operation path, which stays as the reference implementation for it (tests/test_gpu_decode_program.py).
//
// Replaces, per decode step: lia/modeling_opt.py:1379-1491 (layer loop) driving decoder.py:172-335 and
// attentions.py:312-557, plus models.py:423-431 and greedy_search.py:367-395 for the head.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "umma.cuh"
#include "gemm_shared.cuh"
#include "small_ops.cuh"

namespace {

enum { OP_GEMM = 0, OP_LAYERNORM = 1, OP_ATTN = 2, OP_EMBED = 3, OP_ARGMAX = 4 };
constexpr unsigned long long DEP_TIMEOUT_NS = 4000000000ull;

struct alignas(128) DecodeOp {
  CUtensorMap tmA;        // GEMM: weights [N,K], box 128 x 64 (128-byte swizzle).  ATTN: K cache [rows, cache_batch*H*d], box att_rows x d
  CUtensorMap tmB;        // GEMM: activations [M,K], box BN x 64.  ATTN: V cache
  EpiParams ep;           // GEMM: epilogue block (pos0 of a QKV epilogue is replaced by the launch's `pos0`)
  int kind;
  int ncta;               // GEMM: CTAs that share it (a prefix of the grid -- the stand-alone kernel's grid)
  int k_blocks, tiles_a, streamk;
  int tp_index;           // GEMM fused with its all-reduce: ordinal among the program's exchanges (epoch offset), else -1
  // LAYERNORM (x, w, b -> y; rows x h)   EMBED (tok, pos -> y; rows = B, S = 1)   ARGMAX (x = logits [rows, V])
  const bf16* x;
  const bf16* w;
  const bf16* b;
  bf16* y;
  int rows, h;
  float eps;
  // ATTN
  const bf16* q;
  const bf16* kc;
  const bf16* vc;
  int B, H, d, cache_batch, b0, att_rows, att_bytes;
  // EMBED
  const int64_t* mask;
  int mask_ld, vocab, max_pos_rows;
  // ARGMAX
  int V;
};

// sync[0] arrivals recorded before this launch   sync[1] arrival counter   sync[2] CTAs that left   sync[3] error word
__device__ __forceinline__ void wait_arrivals(const int* sync, int target, int* err) {
  unsigned long long t0 = 0;
  unsigned it = 0;
  while ((int)(ld_acquire_gpu(sync + 1) - target) < 0) {
    if ((++it & 1023u) == 0) {
      if (*reinterpret_cast<volatile int*>(err) != 0) return;
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
      if (t0 == 0) t0 = t;
      else if (t - t0 > DEP_TIMEOUT_NS) {          // a CTA that never arrives must not hang the GPU
        atomicExch(err, 1);
        return;
      }
    }
  }
}

// (b,h) pairs of an attention operation this CTA owns, and K (= V) tiles per pair
__device__ __forceinline__ int attn_pairs(const DecodeOp& op, int cta, int ncta) {
  const int total = op.B * op.H;
  return cta < total ? (total - cta + ncta - 1) / ncta : 0;
}
__device__ __forceinline__ int attn_tiles(const DecodeOp& op, int pos0) { return (pos0 + op.att_rows - 1) / op.att_rows; }

template <int BN, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
lia_decode_program_kernel(const DecodeOp* __restrict__ ops, int n_ops, int n_tp_ops, int* __restrict__ tp_ctl, int pos0,
                          const int64_t* __restrict__ ids_in, int64_t* __restrict__ ids_out, int suppress, int* __restrict__ sync,
                          float* __restrict__ ws, int* __restrict__ flags, int dbg_delay_ns, volatile int* __restrict__ progress, unsigned long long* __restrict__ tl) {
  using L = SmemLayout<true, BN, STAGES>;
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
  const int cta = (int)blockIdx.x, G = (int)gridDim.x;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr_smem)),
                 "r"((uint32_t)tmem_cols(BN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();
  int* err = sync + 3;
  // debug timeline (LIA_PROGRAM_DEBUG=1): %globaltimer per operation and role for three CTAs
  const int tl_slot = (cta == 0) ? 0 : (cta == G / 2) ? 1 : (cta == G - 1) ? 2 : -1;
  auto tstamp = [&](int i, int role) {
    if (tl != nullptr && tl_slot >= 0 && i < 4096) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
      tl[((size_t)tl_slot * 4096 + i) * 4 + role] = t;
    }
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      pdl_wait();
      const int base = *reinterpret_cast<volatile int*>(sync);
      int stage = 0;
      uint32_t phase = 0;
      auto advance = [&]() {
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      };
      for (int i = 0; i < n_ops; ++i) {
        const DecodeOp& op = ops[i];
        if (progress) progress[cta * 4 + 0] = i + 1;
        tstamp(i, 0);
        if (op.kind == OP_GEMM) {
          // weight tiles depend on nothing: put up to a ring of them in flight BEFORE the operations this GEMM's
          // activations come from are known to be complete
          int npre = 0, pst = stage;
          uint32_t pph = phase;
          {
            Sched<true> pre(op.k_blocks, op.tiles_a, 1, op.streamk, false, cta, op.ncta);
            Work w;
            while (npre < STAGES && pre.next(w)) {
              for (int kb = w.kb0; kb < w.kb1 && npre < STAGES; ++kb, ++npre) {
                mbar_wait(empty_bar(pst), pph ^ 1u);
                mbar_expect_tx(full_bar(pst), L::STAGE_BYTES);
                tma_load_2d(smem_base + pst * L::STAGE_BYTES, &op.tmA, kb * BLOCK_K, w.ta * TILE_A, full_bar(pst));
                if (++pst == STAGES) {
                  pst = 0;
                  pph ^= 1u;
                }
              }
            }
          }
          if (i > 0) {
            wait_arrivals(sync, base + i * G, err);
            asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy writes of other SMs -> our TMA (async proxy) reads
            if (dbg_delay_ns > 0) __nanosleep(dbg_delay_ns);
          }
          Sched<true> sched(op.k_blocks, op.tiles_a, 1, op.streamk, false, cta, op.ncta);
          Work w;
          int idx = 0;
          while (sched.next(w)) {
            for (int kb = w.kb0; kb < w.kb1; ++kb, ++idx) {
              const uint32_t sa = smem_base + stage * L::STAGE_BYTES;
              const uint32_t sb = sa + L::A_BYTES;
              if (idx < npre) {
                tma_load_2d(sb, &op.tmB, kb * BLOCK_K, 0, full_bar(stage));     // the weight tile is already in flight
              } else {
                mbar_wait(empty_bar(stage), phase ^ 1u);
                mbar_expect_tx(full_bar(stage), L::STAGE_BYTES);
                tma_load_2d(sa, &op.tmA, kb * BLOCK_K, w.ta * TILE_A, full_bar(stage));
                tma_load_2d(sb, &op.tmB, kb * BLOCK_K, 0, full_bar(stage));
              }
              advance();
            }
          }
        } else if (op.kind == OP_ATTN) {
          // cached rows [0, pos0) were written by earlier launches: nothing in this program touches them, so their tiles
          // stream while the QKV projection before this operation is still finishing (the row it appends, pos0, is read
          // by the compute warps straight from global memory after the dependency)
          const int nt = attn_tiles(op, pos0);
          const int total = op.B * op.H;
          for (int pair = cta; pair < total; pair += G) {
            const int b = pair / op.H, hh = pair - b * op.H;
            const int col = ((op.b0 + b) * op.H + hh) * op.d;
            for (int kv = 0; kv < 2; ++kv) {
              for (int t = 0; t < nt; ++t) {
                mbar_wait(empty_bar(stage), phase ^ 1u);
                mbar_expect_tx(full_bar(stage), (uint32_t)op.att_bytes);
                tma_load_2d(smem_base + stage * L::STAGE_BYTES, kv == 0 ? &op.tmA : &op.tmB, col, t * op.att_rows, full_bar(stage));
                advance();
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TILE_A, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int i = 0; i < n_ops; ++i) {
        const DecodeOp& op = ops[i];
        if (progress) progress[cta * 4 + 1] = i + 1;
        tstamp(i, 1);
        if (op.kind == OP_GEMM) {
          Sched<true> sched(op.k_blocks, op.tiles_a, 1, op.streamk, false, cta, op.ncta);
          Work w;
          while (sched.next(w)) {
            mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
            tcgen05_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
            for (int kb = w.kb0; kb < w.kb1; ++kb) {
              mbar_wait(full_bar(stage), phase);
              tcgen05_fence_after();
              const uint32_t sa = smem_base + stage * L::STAGE_BYTES;
              const uint64_t da = make_smem_desc(sa);
              const uint64_t db = make_smem_desc(sa + L::A_BYTES);
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                tcgen05_mma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > w.kb0 || k > 0) ? 1u : 0u);
              tcgen05_commit(empty_bar(stage));
              if (++stage == STAGES) {
                stage = 0;
                phase ^= 1u;
              }
            }
            tcgen05_commit(tfull_bar(acc));
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
          }
        } else if (op.kind == OP_ATTN) {
          // the compute warps consume this operation's ring slots.  This thread must still FOLLOW them one by one: an
          // mbarrier parity wait only tells the current phase from the previous one, so a waiter that skipped a phase of
          // a slot would see its next wait on that slot succeed at once (on the phase it skipped)
          const unsigned n = (unsigned)attn_pairs(op, cta, G) * 2u * (unsigned)attn_tiles(op, pos0);
          for (unsigned k = 0; k < n; ++k) {
            mbar_wait(full_bar(stage), phase);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue / compute warps (128 threads) =====================
    const int ew = warp - EPI_WARP0;
    const int et = threadIdx.x - EPI_WARP0 * 32;
    pdl_wait();
    const int base = *reinterpret_cast<volatile int*>(sync);
    const int epoch0 = (n_tp_ops > 0) ? *reinterpret_cast<volatile int*>(tp_ctl) : 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    int stage = 0;              // ring position (attention tiles are consumed by these warps)
    uint32_t phase = 0;
    float* stgf = reinterpret_cast<float*>(smem_gen + STAGES * L::STAGE_BYTES);   // BN x SWAP_LD floats of scratch
    auto sync128 = [] { epi_bar_sync(); };
    for (int i = 0; i < n_ops; ++i) {
      const DecodeOp& op = ops[i];
      if (progress && et == 0) progress[cta * 4 + 2] = i + 1;
      // every operation starts once all earlier ones are complete on every CTA (also orders re-used buffers)
      if (i > 0) {
        if (et == 0) wait_arrivals(sync, base + i * G, err);
        epi_bar_sync();
      }
      const int kind = op.kind;
      if (progress && et == 0) progress[cta * 4 + 3] = i + 1;
      if (et == 0) tstamp(i, 2);
      if (kind == OP_GEMM) {
        // parameter block -> shared memory once per operation (uniform reads afterwards); a QKV epilogue appends at pos0
        EpiParams* ps = reinterpret_cast<EpiParams*>(smem_gen + L::BAR_OFFSET + (2 * STAGES + 4) * 8 + 16);
        {
          const uint32_t* src = reinterpret_cast<const uint32_t*>(&op.ep);
          uint32_t* dst = reinterpret_cast<uint32_t*>(ps);
          for (int k = et; k < (int)(sizeof(EpiParams) / 4); k += 128) dst[k] = src[k];
          epi_bar_sync();
          if (et == 0 && ps->mode == LIA_EPI_QKV) ps->pos0 = pos0;
          epi_bar_sync();
        }
        const EpiParams& p = *ps;
        SwapEpiCtx ctx;
        ctx.stgf = stgf;
        ctx.ws = ws;
        ctx.flags = flags;
        ctx.trace = nullptr;
        ctx.k_blocks = op.k_blocks;
        ctx.tiles_a = op.tiles_a;
        ctx.cta = cta;
        ctx.ncta = op.ncta;
        ctx.ew = ew;
        ctx.lane = lane;
        ctx.et = et;
        ctx.tp_on = (p.mode == EPI_TP);
        ctx.epoch = epoch0 + 1 + (op.tp_index > 0 ? op.tp_index : 0);
        ctx.parity = ctx.epoch & 1;
        ctx.tp_err = ctx.tp_on ? p.tp.ctl(p.tp.rank) + 2 : nullptr;
        Sched<true> sched(op.k_blocks, op.tiles_a, 1, op.streamk, false, cta, op.ncta);
        {   // ring slots of this CTA's span (consumed by the MMA warp)
          const unsigned s2 = (unsigned)stage + (sched.end - sched.pos);
          if ((s2 / STAGES) & 1u) phase ^= 1u;
          stage = (int)(s2 % STAGES);
        }
        Work w;
        while (sched.next(w)) {
          constexpr int ITERS = BN / 8;
          const int c8 = et & 15, m0 = et >> 4;
          const int rows = min(BN, p.M);
          const int n_col = w.ta * TILE_A + c8 * 8;
          float biasf[8];
          uint4 resv[ITERS];
          if (w.kb0 == 0 && n_col < p.N) {
            epilogue_load_bias8(p, n_col, biasf);
            const bool need_res = !(ctx.tp_on && (p.tp.opts & 128));
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
              const int m = m0 + it * 8;
              resv[it] = (m < rows && need_res) ? epilogue_load_residual8(p, m, n_col) : make_uint4(0, 0, 0, 0);
            }
          }
          mbar_wait(tfull_bar(acc), acc_phase);
          tcgen05_fence_after();
          const uint32_t taddr = tmem_base + (uint32_t)(acc * BN) + ((uint32_t)(ew * 32) << 16);
          swap_epilogue_tile<BN>(p, ctx, w, taddr, tempty_bar(acc), biasf, resv);
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1u;
        }
      } else if (kind == OP_LAYERNORM) {
        if (op.h <= 2048) {
          for (int r0 = cta * 4; r0 < op.rows; r0 += G * 4)
            layernorm_rows<32, 8>(op.x, op.w, op.b, op.y, op.rows, op.h, op.eps, r0, et, stgf, sync128);
        } else {
          for (int r0 = cta; r0 < op.rows; r0 += G) {
            layernorm_rows<128, 16>(op.x, op.w, op.b, op.y, op.rows, op.h, op.eps, r0, et, stgf, sync128);
            epi_bar_sync();                               // `red` scratch is re-used by the next row
          }
        }
      } else if (kind == OP_EMBED) {
        for (int r = cta; r < op.rows; r += G) {
          embed_row(ids_in, op.w, op.b, op.y, r, 1, op.h, pos0, op.vocab, op.max_pos_rows, op.mask, op.mask_ld, et,
                    reinterpret_cast<long long*>(stgf), sync128);
          epi_bar_sync();
        }
      } else if (kind == OP_ARGMAX) {
        for (int r = cta; r < op.rows; r += G) {
          argmax_row<128>(op.x + (size_t)r * op.V, ids_out + r, op.V, suppress, et, stgf, reinterpret_cast<int*>(stgf + 8), sync128);
          epi_bar_sync();
        }
      } else if (kind == OP_ATTN) {
        // attn_decode.cu's arithmetic, same order (splits == 1): s = bf16(q.k), p = bf16(exp(s - m) / l), ctx = bf16(sum p v);
        // keys [0, pos0) come through the ring as tiles of att_rows rows, key pos0 (this step's) straight from the cache
        const int D = op.d;
        const int LPR = D / 8, RPW = 32 / LPR;            // lanes per cached row, rows per warp pass
        const int chunk = lane % LPR, rsub = lane / LPR;
        const int T = pos0 + 1;
        const int nt = attn_tiles(op, pos0);
        const int total = op.B * op.H;
        float* sc = stgf;                                 // [T] scores -> probabilities
        float* red = stgf + BN * SWAP_LD - 8 - 4 * 128;   // 4 floats (+4 spare), then accs[4][128]
        float* accs = red + 8;
        const size_t row_stride = (size_t)op.cache_batch * op.H * D;
        for (int pair = cta; pair < total; pair += G) {
          const int b = pair / op.H, hh = pair - b * op.H;
          float qf[8];
          unpack8(ldg_act(op.q + (size_t)pair * D + chunk * 8), qf);
          const size_t gbase = ((size_t)(op.b0 + b) * op.H + hh) * D + chunk * 8;
          const int slot = ew * RPW + rsub;               // this lane group owns keys slot, slot + 4*RPW, ...
          // ---- phase A: scores
          for (int t = 0; t < nt; ++t) {
            mbar_wait(full_bar(stage), phase);
            const uint8_t* tile = smem_gen + stage * L::STAGE_BYTES;
            const int t0 = t * op.att_rows;
            const int n = min(op.att_rows, pos0 - t0);    // old keys in this tile
            for (int rb = 0; rb < n; rb += 4 * RPW) {       // uniform trip count: the shuffles below need every lane
              const int r = rb + slot;
              float dot = 0.f;
              if (r < n) {
                float kf[8];
                unpack8(*reinterpret_cast<const uint4*>(tile + (size_t)r * D * 2 + chunk * 16), kf);
#pragma unroll
                for (int e = 0; e < 8; ++e) dot = fmaf(qf[e], kf[e], dot);
              }
              for (int o = LPR / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
              if (chunk == 0 && r < n) sc[t0 + r] = bf16r(dot);
            }
            epi_bar_sync();                               // the tile has been read by all four warps
            if (et == 0) mbar_arrive(empty_bar(stage));
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
          {                                               // this step's key: the row the QKV epilogue just appended
            float kf[8];
            unpack8(ldg_act(op.kc + gbase + (size_t)pos0 * row_stride), kf);
            float dot = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) dot = fmaf(qf[e], kf[e], dot);
            for (int o = LPR / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
            if (chunk == 0 && (pos0 % (4 * RPW)) == slot) sc[pos0] = bf16r(dot);   // every lane group computes it, its owner keeps it
          }
          epi_bar_sync();
          // ---- softmax over [0, T)
          float m = -INFINITY;
          for (int k = et; k < T; k += 128) m = fmaxf(m, sc[k]);
          m = warp_max(m);
          if (lane == 0) red[ew] = m;
          epi_bar_sync();
          m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
          epi_bar_sync();
          float l = 0.f;
          for (int k = et; k < T; k += 128) {
            const float e = expf(sc[k] - m);
            sc[k] = e;
            l += e;
          }
          l = warp_sum(l);
          if (lane == 0) red[ew] = l;
          epi_bar_sync();
          l = red[0] + red[1] + red[2] + red[3];
          for (int k = et; k < T; k += 128) sc[k] = bf16r(sc[k] / l);
          epi_bar_sync();
          // ---- phase B: ctx = sum_t p_t v_t
          float acc8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) acc8[e] = 0.f;
          for (int t = 0; t < nt; ++t) {
            mbar_wait(full_bar(stage), phase);
            const uint8_t* tile = smem_gen + stage * L::STAGE_BYTES;
            const int t0 = t * op.att_rows;
            const int n = min(op.att_rows, pos0 - t0);
            for (int r = slot; r < n; r += 4 * RPW) {
              const float pj = sc[t0 + r];
              float vf[8];
              unpack8(*reinterpret_cast<const uint4*>(tile + (size_t)r * D * 2 + chunk * 16), vf);
#pragma unroll
              for (int e = 0; e < 8; ++e) acc8[e] = fmaf(pj, vf[e], acc8[e]);
            }
            epi_bar_sync();
            if (et == 0) mbar_arrive(empty_bar(stage));
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
          if ((pos0 % (4 * RPW)) == slot) {
            const float pj = sc[pos0];
            float vf[8];
            unpack8(ldg_act(op.vc + gbase + (size_t)pos0 * row_stride), vf);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc8[e] = fmaf(pj, vf[e], acc8[e]);
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            for (int o = 16; o >= LPR; o >>= 1) acc8[e] += __shfl_xor_sync(0xffffffffu, acc8[e], o);
          }
          if (rsub == 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) accs[ew * 128 + chunk * 8 + e] = acc8[e];
          }
          epi_bar_sync();
          if (et < D) {
            const float o = accs[et] + accs[128 + et] + accs[256 + et] + accs[384 + et];
            op.y[(size_t)pair * D + et] = __float2bfloat16_rn(o);
          }
          epi_bar_sync();                                 // sc / accs are re-used by the next pair
        }
      }
      // this CTA's part of operation i is complete and visible: every thread's writes are fenced, the CTA meets, and one
      // thread publishes the arrival with release semantics
      __threadfence();
      epi_bar_sync();
      if (et == 0) {
        tstamp(i, 3);
        __threadfence();
        asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(sync + 1) : "memory");
      }
    }
    // leave: the last CTA publishes the counters for the next launch (every CTA read them on entry)
    if (et == 0) {
      __threadfence();
      if (atomicAdd(sync + 2, 1) == G - 1) {
        sync[2] = 0;
        if (n_tp_ops > 0) {
          tp_ctl[1] = 0;
          *reinterpret_cast<volatile int*>(tp_ctl) = epoch0 + n_tp_ops;
        }
        __threadfence();
        *reinterpret_cast<volatile int*>(sync) = base + n_ops * G;
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)tmem_cols(BN)) : "memory");
  }
}

// 2-D bf16 tensor [rows, cols] (row pitch = cols), box = [box_rows, box_cols] WITHOUT swizzle (attention tiles are read by
// ordinary shared-memory loads); out-of-bounds elements read as zeros
int make_tmap_plain(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, int box_rows, int box_cols) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    lia_set_error("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    return LIA_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    lia_set_error("cuTensorMapEncodeTiled (plain) failed with CUresult %d (rows=%llu cols=%llu box=%dx%d)", (int)r,
                  (unsigned long long)rows, (unsigned long long)cols, box_rows, box_cols);
    return LIA_ERR_CUDA;
  }
  return LIA_OK;
}

int bn_for_rows(int M) { return M <= 16 ? 16 : M <= 32 ? 32 : M <= 64 ? 64 : 128; }
int stage_bytes_for_bn(int bn) { return TILE_A * BLOCK_K * 2 + bn * BLOCK_K * 2; }

}  // namespace

// debug only (LIA_PROGRAM_DEBUG=1): per-CTA progress words in mapped host memory, readable while a launch hangs
static int* g_progress_host = nullptr;
static int* g_progress_dev = nullptr;
extern "C" const int* lia_debug_program_progress(void) {
  if (!g_progress_host) {
    const char* e = getenv("LIA_PROGRAM_DEBUG");
    if (e && atoi(e) != 0 && cudaHostAlloc((void**)&g_progress_host, 4096 * sizeof(int) + 3 * 4096 * 4 * 8, cudaHostAllocMapped) == cudaSuccess) {
      memset(g_progress_host, 0, 4096 * sizeof(int) + 3 * 4096 * 4 * 8);
      cudaHostGetDevicePointer((void**)&g_progress_dev, g_progress_host, 0);
    }
  }
  return g_progress_host;
}

struct LiaProgram {
  int M = 0;                       // rows (tokens) every operation works on: fixes the kernel variant
  int bn = 0;
  std::vector<DecodeOp> ops;
  DecodeOp* d_ops = nullptr;
  int* d_sync = nullptr;           // 4 ints
  void* d_ws = nullptr;            // stream-K flags + pieces (same layout as the stand-alone GEMM's workspace)
  size_t ws_bytes = 0;
  int n_tp = 0;
  int* tp_ctl = nullptr;
  bool finalized = false;
};

extern "C" LiaProgram* lia_program_create(int rows) {
  if (rows <= 0 || rows > 128) {
    lia_set_error("lia_program_create: rows must be in [1, 128] (got %d): programs are for decode-shaped chains", rows);
    return nullptr;
  }
  LiaProgram* p = new (std::nothrow) LiaProgram();
  if (!p) {
    lia_set_error("lia_program_create: out of host memory");
    return nullptr;
  }
  p->M = rows;
  p->bn = bn_for_rows(rows);
  return p;
}

static int program_check(LiaProgram* p, const char* fn) {
  LIA_CHECK_ARG(p != nullptr, "%s: null program", fn);
  LIA_CHECK_ARG(!p->finalized, "%s: the program is already finalized", fn);
  return LIA_OK;
}

extern "C" int lia_program_add_gemm(LiaProgram* p, const void* A, const void* W, const void* bias, const void* residual, void* out,
                                    int M, int N, int K, int epilogue, const LiaQkvArgs* qkv, const LiaTpArgs* tp) {
  int rc = program_check(p, "lia_program_add_gemm");
  if (rc != LIA_OK) return rc;
  LIA_CHECK_ARG(M == p->M, "lia_program_add_gemm: M=%d but the program was created for %d rows", M, p->M);
  LIA_CHECK_ARG(tp != nullptr || (epilogue >= LIA_EPI_BIAS && epilogue <= LIA_EPI_QKV), "lia_program_add_gemm: unknown epilogue %d", epilogue);
  DecodeOp op;
  memset(&op, 0, sizeof(op));
  Plan pl;
  rc = gemm_fill_params(A, W, bias, residual, out, M, N, K, tp ? EPI_TP : epilogue, qkv, tp, op.ep, pl);
  if (rc != LIA_OK) return rc;
  LIA_CHECK_ARG(pl.swap && pl.bn == p->bn, "lia_program_add_gemm: unexpected plan");
  if (pl.streamk && pl.grid * (int)sizeof(int) > COUNTER_BYTES) {
    pl.streamk = 0;
    if (pl.grid > pl.tiles_a) pl.grid = pl.tiles_a;
  }
  op.kind = OP_GEMM;
  op.ncta = pl.grid;
  op.k_blocks = pl.k_blocks;
  op.tiles_a = pl.tiles_a;
  op.streamk = pl.streamk;
  op.tp_index = -1;
  if (tp != nullptr) {
    int* ctl = reinterpret_cast<int*>(reinterpret_cast<char*>(tp->arena[tp->rank]) + tp->ctl_off);
    LIA_CHECK_ARG(p->tp_ctl == nullptr || p->tp_ctl == ctl, "lia_program_add_gemm: all exchanges of a program must use one arena");
    p->tp_ctl = ctl;
    op.tp_index = p->n_tp++;
  }
  if ((rc = make_tmap(&op.tmA, W, N, K, TILE_A)) != LIA_OK) return rc;
  if ((rc = make_tmap(&op.tmB, A, M, K, pl.bn)) != LIA_OK) return rc;
  const size_t need = plan_workspace(pl);
  if (need > p->ws_bytes) p->ws_bytes = need;
  p->ops.push_back(op);
  return LIA_OK;
}

extern "C" int lia_program_add_layernorm(LiaProgram* p, const void* x, const void* w, const void* b, void* y, int rows, int h, float eps) {
  int rc = program_check(p, "lia_program_add_layernorm");
  if (rc != LIA_OK) return rc;
  LIA_CHECK_ARG(x && w && b && y, "lia_program_add_layernorm: null pointer");
  LIA_CHECK_ARG(rows > 0 && h > 0 && h % 8 == 0 && h <= 16384, "lia_program_add_layernorm: need h %% 8 == 0 and h <= 16384 (h=%d)", h);
  DecodeOp op;
  memset(&op, 0, sizeof(op));
  op.kind = OP_LAYERNORM;
  op.x = reinterpret_cast<const bf16*>(x);
  op.w = reinterpret_cast<const bf16*>(w);
  op.b = reinterpret_cast<const bf16*>(b);
  op.y = reinterpret_cast<bf16*>(y);
  op.rows = rows;
  op.h = h;
  op.eps = eps;
  p->ops.push_back(op);
  return LIA_OK;
}

extern "C" int lia_program_add_attn_decode(LiaProgram* p, const void* q, const void* k_cache, const void* v_cache, void* out, int B,
                                           int H, int d, int cache_batch, int b0, int cache_rows) {
  int rc = program_check(p, "lia_program_add_attn_decode");
  if (rc != LIA_OK) return rc;
  LIA_CHECK_ARG(q && k_cache && v_cache && out, "lia_program_add_attn_decode: null pointer");
  LIA_CHECK_ARG(d == 64 || d == 128, "lia_program_add_attn_decode: head_dim must be 64 or 128 (got %d)", d);
  LIA_CHECK_ARG(B > 0 && H > 0 && cache_rows > 0, "lia_program_add_attn_decode: B, H, cache_rows must be positive");
  LIA_CHECK_ARG(b0 >= 0 && b0 + B <= cache_batch, "lia_program_add_attn_decode: batch window [%d,%d) outside cache batch %d", b0, b0 + B, cache_batch);
  // scores of all cached positions live in the GEMM's staging area
  const int max_T = p->bn * SWAP_LD - 8 - 4 * 128;
  LIA_CHECK_ARG(cache_rows <= max_T, "lia_program_add_attn_decode: %d cache rows exceed the %d scores a %d-row program can hold", cache_rows,
                max_T, p->M);
  DecodeOp op;
  memset(&op, 0, sizeof(op));
  op.kind = OP_ATTN;
  op.q = reinterpret_cast<const bf16*>(q);
  op.kc = reinterpret_cast<const bf16*>(k_cache);
  op.vc = reinterpret_cast<const bf16*>(v_cache);
  op.y = reinterpret_cast<bf16*>(out);
  op.B = B; op.H = H; op.d = d; op.cache_batch = cache_batch; op.b0 = b0;
  int rows = stage_bytes_for_bn(p->bn) / (d * 2);
  if (rows > 256) rows = 256;
  rows -= rows % 8;                 // whole passes of the 4 warps x RPW lane groups
  op.att_rows = rows;
  op.att_bytes = rows * d * 2;
  if ((rc = make_tmap_plain(&op.tmA, k_cache, (uint64_t)cache_rows, (uint64_t)cache_batch * H * d, rows, d)) != LIA_OK) return rc;
  if ((rc = make_tmap_plain(&op.tmB, v_cache, (uint64_t)cache_rows, (uint64_t)cache_batch * H * d, rows, d)) != LIA_OK) return rc;
  p->ops.push_back(op);
  return LIA_OK;
}

extern "C" int lia_program_add_embed(LiaProgram* p, const int64_t* attention_mask, int mask_ld, const void* embed_tokens,
                                     const void* embed_positions, void* out, int B, int h, int vocab, int max_pos_rows) {
  int rc = program_check(p, "lia_program_add_embed");
  if (rc != LIA_OK) return rc;
  LIA_CHECK_ARG(out && (embed_tokens || embed_positions), "lia_program_add_embed: null pointer");
  LIA_CHECK_ARG(B > 0 && h > 0 && h % 8 == 0, "lia_program_add_embed: bad shape");
  DecodeOp op;
  memset(&op, 0, sizeof(op));
  op.kind = OP_EMBED;
  op.w = reinterpret_cast<const bf16*>(embed_tokens);
  op.b = reinterpret_cast<const bf16*>(embed_positions);
  op.y = reinterpret_cast<bf16*>(out);
  op.rows = B;
  op.h = h;
  op.mask = attention_mask;
  op.mask_ld = mask_ld;
  op.vocab = vocab;
  op.max_pos_rows = max_pos_rows;
  p->ops.push_back(op);
  return LIA_OK;
}

extern "C" int lia_program_add_argmax(LiaProgram* p, const void* logits, int B, int V) {
  int rc = program_check(p, "lia_program_add_argmax");
  if (rc != LIA_OK) return rc;
  LIA_CHECK_ARG(logits && B > 0 && V > 0, "lia_program_add_argmax: bad arguments");
  LIA_CHECK_ARG(((size_t)V * 2) % 16 == 0 || B == 1, "lia_program_add_argmax: V*2 must be a multiple of 16 bytes for B > 1 (V=%d)", V);
  DecodeOp op;
  memset(&op, 0, sizeof(op));
  op.kind = OP_ARGMAX;
  op.x = reinterpret_cast<const bf16*>(logits);
  op.rows = B;
  op.V = V;
  p->ops.push_back(op);
  return LIA_OK;
}

extern "C" int lia_program_finalize(LiaProgram* p) {
  int rc = program_check(p, "lia_program_finalize");
  if (rc != LIA_OK) return rc;
  LIA_CHECK_ARG(!p->ops.empty(), "lia_program_finalize: empty program");
  if (p->ws_bytes < (size_t)COUNTER_BYTES) p->ws_bytes = COUNTER_BYTES;
  LIA_CUDA(cudaMalloc(&p->d_ops, p->ops.size() * sizeof(DecodeOp)));
  LIA_CUDA(cudaMemcpy(p->d_ops, p->ops.data(), p->ops.size() * sizeof(DecodeOp), cudaMemcpyHostToDevice));
  LIA_CUDA(cudaMalloc(&p->d_sync, 16 * sizeof(int)));
  LIA_CUDA(cudaMemset(p->d_sync, 0, 16 * sizeof(int)));
  LIA_CUDA(cudaMalloc(&p->d_ws, p->ws_bytes));
  LIA_CUDA(cudaMemset(p->d_ws, 0, p->ws_bytes));
  LIA_CUDA(cudaDeviceSynchronize());
  p->finalized = true;
  return LIA_OK;
}

template <int BN, int STAGES>
static int program_launch(LiaProgram* p, int pos0, const int64_t* ids_in, int64_t* ids_out, int suppress, cudaStream_t stream) {
  using L = SmemLayout<true, BN, STAGES>;
  constexpr int SMEM = L::TOTAL + 512;      // + the epilogue parameter block of the running operation
  static_assert(sizeof(EpiParams) <= 480, "parameter block must fit its shared-memory slot");
  static_assert(SMEM <= 232448, "shared memory budget exceeded");
  auto kern = lia_decode_program_kernel<BN, STAGES>;
  static bool configured = false;
  if (!configured) {
    LIA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  int* flags = reinterpret_cast<int*>(p->d_ws);
  float* ws = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(p->d_ws) + COUNTER_BYTES);
  const char* de = getenv("LIA_PROGRAM_DEBUG_DELAY_NS");   // debug probe: delay between a dependency wait and the TMA loads after it
  const int dbg = de ? atoi(de) : 0;
  LIA_CUDA(lia_launch(kern, dim3(lia_sm_count()), dim3(NUM_THREADS), SMEM, stream, (const DecodeOp*)p->d_ops, (int)p->ops.size(), p->n_tp,
                      p->tp_ctl, pos0, ids_in, ids_out, suppress, p->d_sync, ws, flags, dbg, (volatile int*)g_progress_dev,
                      g_progress_dev ? reinterpret_cast<unsigned long long*>(g_progress_dev + 4096) : (unsigned long long*)nullptr));
  return LIA_OK;
}

extern "C" int lia_program_run(LiaProgram* p, int pos0, const int64_t* ids_in, int64_t* ids_out, int suppress_id, lia_stream_t stream_) {
  LIA_CHECK_ARG(p != nullptr && p->finalized, "lia_program_run: the program is not finalized");
  LIA_CHECK_ARG(pos0 >= 0, "lia_program_run: negative position");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  switch (p->bn) {
    case 16: return program_launch<16, 8>(p, pos0, ids_in, ids_out, suppress_id, stream);
    case 32: return program_launch<32, 8>(p, pos0, ids_in, ids_out, suppress_id, stream);
    case 64: return program_launch<64, 8>(p, pos0, ids_in, ids_out, suppress_id, stream);
    default: return program_launch<128, 4>(p, pos0, ids_in, ids_out, suppress_id, stream);
  }
}

extern "C" int lia_program_error(LiaProgram* p) {
  LIA_CHECK_ARG(p != nullptr && p->finalized, "lia_program_error: the program is not finalized");
  LIA_CUDA(cudaDeviceSynchronize());
  int v = 0;
  LIA_CUDA(cudaMemcpy(&v, p->d_sync + 3, sizeof(int), cudaMemcpyDeviceToHost));
  if (v != 0) {
    LIA_CUDA(cudaMemset(p->d_sync + 3, 0, sizeof(int)));
    lia_set_error("decode program: a CTA timed out waiting for an earlier operation (error word %d)", v);
    return 1;
  }
  return 0;
}

extern "C" int lia_program_num_ops(LiaProgram* p) { return p ? (int)p->ops.size() : 0; }

extern "C" int lia_program_destroy(LiaProgram* p) {
  if (!p) return LIA_OK;
  if (p->d_ops) cudaFree(p->d_ops);
  if (p->d_sync) cudaFree(p->d_sync);
  if (p->d_ws) cudaFree(p->d_ws);
  delete p;
  return LIA_OK;
}
