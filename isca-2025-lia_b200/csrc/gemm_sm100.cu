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
#include "gemm_shared.cuh"

namespace {
// ------------------------------------------------------------------ the kernel
template <bool SWAP, int BN, int STAGES, bool TP, bool PAIR = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
lia_gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const __grid_constant__ EpiParams p,
                        int k_blocks, int streamk, int tiles_a, int tiles_b, float* __restrict__ ws,
                        int* __restrict__ flags, unsigned long long* __restrict__ trace) {
  static_assert(!PAIR || (!SWAP && !TP && BN == 256), "CTA pairs: plain prefill projections with 256-wide tiles only");
  using L = SmemLayout<SWAP, BN, STAGES, PAIR>;
  constexpr int ACC_COLS = BN;   // TMEM columns from one accumulator to the other
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
        // The CTA barrier orders every epilogue thread's stores of the covered tile before the signalling threads, whose
        // st.release.sys is the release pattern that publishes them (release is cumulative over what the barrier made
        // visible to the storing thread).  Round 1 had all 128 threads execute a system-scope fence here first: ~8 us per
        // tile, 100 us per launch of a short-K projection where no MMA time hides it (scripts/tp_microbench.py, TP2
        // out_proj prefill: 509 -> 405 us).
        epi_bar_sync();
        if (et < tp.world) {
          if (sig_data_u >= 0 && et == 0) st_release_sys(tp.data_flag(sig_data_u % tp.world, sig_data_u, tp.rank), epoch);
          if (sig_done_u >= 0 && et != tp.rank) st_release_sys(tp.done_flag(et, sig_done_u), epoch);
        }
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
      if (tp.mc != nullptr) {
        // NVLS: every rank's partial of tile u sits at the same offset of its OWN arena; one multimem.ld_reduce returns the
        // sum of the `world` copies (reduced inside the NVLink switch, fp32 accumulation, one bf16 rounding = the
        // reference's message dtype), one multimem.st delivers the final chunk to every rank's `out`
        const char* mpart = tp.mc + tp.recv_off + (unsigned long long)parity * tp.recv_bytes + (size_t)u * (TILE_A * BN) * 2;
#pragma unroll 1
        for (int c0 = et; c0 < TILE_A * CPR; c0 += 128 * CH) {
          uint4 sumv[CH], res[CH];
          bool ok[CH];
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const int c = c0 + j * 128;
            const int row = c / CPR, ch = c - row * CPR;
            const int m = ta * TILE_A + row, n = tb * BN + ch * 8;
            ok[j] = (m < p.M && n < p.N);
            res[j] = ok[j] ? ldg_act(p.residual + (size_t)m * p.N + n) : make_uint4(0, 0, 0, 0);
            sumv[j] = ok[j] ? multimem_ld_reduce_bf16x8(mpart + ((size_t)row * BN + ch * 8) * 2) : make_uint4(0, 0, 0, 0);
          }
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            if (ok[j]) {
              const int c = c0 + j * 128;
              const int row = c / CPR, ch = c - row * CPR;
              const int m = ta * TILE_A + row, n = tb * BN + ch * 8;
              float f[8], r[8];
              unpack8(sumv[j], f);
              unpack8(res[j], r);
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] = r[i] + f[i];
              multimem_st_bf16x8(tp.mc + tp.out_off + ((size_t)m * p.N + n) * 2, pack8(f));
            }
          }
        }
        sig_done_u = u;
        return;
      }
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
        // Everything a 64-column slab's stores need besides the accumulator -- the bias chunk of this thread's 8 columns and
        // the residual chunks of its 8 rows -- is requested ONE SLAB AHEAD: 8 independent 16-byte loads per thread in flight
        // under the previous slab's TMEM read, transposition and stores.  (Fetched one by one inside the store loop, every
        // store waited out a full L2 round trip: the K = h/8 row-parallel projections of a TP8 prefill ran 4x slower than
        // their MMAs.)
        const int ch_s = lane & 7;
        const bool want_res = !tp_on && p.mode == LIA_EPI_BIAS_RESIDUAL;
        float bias_s[8], bias_nx[8];
        uint4 res_s[8], res_nx[8];
        auto slab_fetch = [&](int c0f, float* bo, uint4* ro) {
          const int n_f = w.tb * BN + c0f + ch_s * 8;
          const bool ok = c0f < BN && n_f < p.N && (BN % 64 == 0 || c0f + ch_s * 8 < BN);
          if (ok) epilogue_load_bias8(p, n_f, bo);
          if (want_res) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int m_f = (w.ta * row_tile + (int)cta_rank) * TILE_A + ew * 32 + it * 4 + (lane >> 3);
              ro[it] = (ok && m_f < p.M) ? ldg_act(p.residual + (size_t)m_f * p.N + n_f) : make_uint4(0, 0, 0, 0);
            }
          }
        };
        slab_fetch(0, bias_nx, res_nx);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 64) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            bias_s[i] = bias_nx[i];
            res_s[i] = res_nx[i];
          }
          slab_fetch(c0 + 64, bias_nx, res_nx);           // the next slab's operands travel while this one is processed
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
#pragma unroll
                  for (int i = 0; i < 8; ++i) f[i] = bf16r(f[i] + bias_s[i]);
                }
                const int u = w.ta * tiles_b + w.tb;
                // peer-store exchange: straight into the owner's receive area, slot [tile][this rank];
                // NVLS exchange: into OUR OWN arena at the tile's offset (the switch reads it from there)
                bf16* dst = (tp.mc != nullptr ? tp.recv(tp.rank, parity) + (size_t)u * (TILE_A * BN)
                                              : tp.recv(u % tp.world, parity) + ((size_t)(u / tp.world) * tp.world + tp.rank) * (TILE_A * BN)) +
                            (size_t)(ew * 32 + r) * BN + c0 + ch * 8;
                *reinterpret_cast<uint4*>(dst) = pack8(f);
              } else {
                epilogue_finish8(p, m, n, f, bias_s, res_s[it]);
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
            if (tp.mc != nullptr && et == 0) __threadfence_system();   // NVLS: the switch reads our own partial from memory too
            // defer the reduction by one owned tile (= `world` tiles of MMA work) so the peers' partials are
            // normally already here and the epilogue warps never stall the tensor pipe on NVLink latency
            if (pend_u >= 0) tp_reduce_owned(pend_u, pend_ta, pend_tb);
            pend_u = u; pend_ta = w.ta; pend_tb = w.tb;
          }
        }
      } else {
        // tile rows = output features (TMEM lanes), columns = tokens: transpose through smem / workspace (gemm_shared.cuh)
        SwapEpiCtx ctx;
        ctx.stgf = reinterpret_cast<float*>(smem_gen + STAGES * L::STAGE_BYTES);
        ctx.ws = ws; ctx.flags = flags; ctx.trace = trace;
        ctx.k_blocks = k_blocks; ctx.tiles_a = tiles_a;
        ctx.cta = (int)blockIdx.x; ctx.ncta = (int)gridDim.x;
        ctx.ew = ew; ctx.lane = lane; ctx.et = et;
        ctx.tp_on = tp_on; ctx.epoch = epoch; ctx.parity = parity; ctx.tp_err = tp_err;
        if constexpr (SWAP) swap_epilogue_tile<BN>(p, ctx, w, taddr, tempty_bar(acc), biasf, resv);
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
  EpiParams ep;
  Plan pl;
  const int rc0 = gemm_fill_params(A, W, bias, residual, out, M, N, K, epilogue, qkv, tp, ep, pl);
  if (rc0 != LIA_OK) return rc0;
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
    if (pl.pair) return launch_pair<256>(pl, tmA, tmB, ep, stream);
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
