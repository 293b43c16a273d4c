/*
 * lia_b200.h -- C ABI of libliab200.so: the B200 (sm_100a) kernels behind LIA's OPT
 * decoder-layer hot path.
 *
 * Boundary rules:
 *   - plain pointers and sizes only; every device pointer is owned by the caller
 *     (PyTorch on the Python side) -- the library never allocates or frees device
 *     memory behind the caller's back;  the only allocations it owns are the ones made
 *     explicitly through the functions below: pinned host arenas, streamer handles and
 *     the peer-mappable device arenas of the tensor-parallel path (lia_p2p_*).
 *   - all work is enqueued on the stream that is passed in; no hidden
 *     synchronisation (contrast lia/modeling_opt.py:1339 torch.cuda.synchronize()).
 *   - no C++ exceptions cross the ABI.  Functions return 0 on success, a negative
 *     LIA_ERR_* code on failure; lia_last_error() returns the message of the last
 *     failure on the calling thread.  (The reference's only C ABI,
 *     lia/cxl/numa_alloc.c:29-33,47-55, signals failure with NULL + stderr; its Python
 *     side raises MemoryError, lia/modeling_opt.py:175.)
 *   - bf16 storage, fp32 accumulation everywhere; rounding points mirror the
 *     reference's eager op sequence.
 *
 * Each entry point names the reference interface it replaces (paths relative to the
 * reference root;  D = intel_extension_for_pytorch/transformers/models/reference/modules/decoder.py,
 * A = .../modules/attentions.py, M = lia/modeling_opt.py).
 */
#ifndef LIA_B200_H_
#define LIA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LIA_ABI_VERSION 5

typedef void* lia_stream_t; /* cudaStream_t */

enum {
  LIA_OK = 0,
  LIA_ERR_INVALID = -1, /* bad argument / unsupported shape */
  LIA_ERR_CUDA = -2,    /* a CUDA runtime/driver call failed */
  LIA_ERR_NOMEM = -3,   /* pinned host allocation failed */
  LIA_ERR_ARCH = -4     /* device is not sm_100 */
};

/* GEMM epilogues.  All start from r1 = bf16(acc_fp32), r2 = bf16(r1 + bias)  (two
 * roundings, as `torch.matmul(x, w.t()) + b` produces: A:393-394,418; D:62,81,88,94). */
enum {
  LIA_EPI_BIAS = 0,          /* out = r2                                   q/k/v style, D:81 */
  LIA_EPI_BIAS_RELU = 1,     /* out = relu(r2)                             fc1, D:92-105 */
  LIA_EPI_BIAS_RESIDUAL = 2, /* out = bf16(residual + r2)                  out_proj/fc2 + D:229,310 */
  LIA_EPI_QKV = 3            /* fused q|k|v projection, see LiaQkvArgs     A:376-418,456-491 */
};

/* Fused QKV projection epilogue.  W is [3*hq, K] = rows of Wq, then Wk, then Wv (hq = local
 * attention width, h or h/TP);  A is [B*S, K] with row m = b*S + s.
 *   columns [0,hq)    -> q_out[m, c]                = bf16(r2 * q_scale)        (A:418, A:456)
 *   columns [hq,2hq)  -> k_cache[pos0+s, b0+b, c]   = r2                        (A:393, A:457-491)
 *   columns [2hq,3hq) -> v_cache[pos0+s, b0+b, c]   = r2                        (A:394)
 * caches are time-major [Tmax, cache_batch, hq] (the reference's [(S+new),B,H,d], A:471-472). */
typedef struct LiaQkvArgs {
  void* q_out;
  void* k_cache;
  void* v_cache;
  int32_t hq;
  int32_t S;           /* tokens per sequence in this call (1 in decode) */
  int32_t pos0;        /* first cache row to write (= tokens already cached) */
  int32_t cache_batch; /* batch dimension of the cache tensors */
  int32_t b0;          /* batch offset of this minibatch inside the cache */
  float q_scale;       /* head_dim ** -0.5, M:413 */
} LiaQkvArgs;

int lia_abi_version(void);
const char* lia_last_error(void);
/* sm count / compute capability of the current device. */
int lia_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* LayerNorm over the last dim: bf16 in, fp32 statistics, bf16 out.
 * Replaces gpu_ln_compute_self_attn / gpu_ln_compute_final (D:107-119) and the decoder's
 * final_layer_norm (M:1563-1564).  h % 8 == 0, h <= 16384. */
int lia_layernorm_bf16(const void* x, const void* w, const void* b, void* y, int rows, int h, float eps,
                       lia_stream_t stream);

/* out[M,N] = epilogue(A[M,K] . W[N,K]^T): tcgen05/TMEM GEMM fed by TMA.
 * Replaces gpu_linear_compute* / gpu_linear_relu_compute* (D:79-105) and the three
 * projections of A:376-418.  A, W, bias, residual, out are bf16 row-major and contiguous
 * (K % 8 == 0, N % 8 == 0, all base pointers 16-byte aligned).
 * `workspace` is used when the kernel splits K (small M): lia_gemm_workspace_bytes() bytes,
 * zero-initialised ONCE by the caller (the kernel leaves it zeroed again). */
size_t lia_gemm_workspace_bytes(int M, int N, int K);
int lia_gemm_bf16(const void* A, const void* W, const void* bias, const void* residual, void* out, int M, int N,
                  int K, int epilogue, const LiaQkvArgs* qkv, void* workspace, size_t workspace_bytes,
                  lia_stream_t stream);

/* Causal self-attention over a prompt (prefill), replaces A:444-449,493-536 for tgt_len > 1.
 * q [B,S,H,d] (already scaled), k/v caches [Tmax, cache_batch, H, d] holding rows [0,S) of
 * sequences b0..b0+B-1;  out [B,S,H*d].  Scores are rounded to bf16, softmax in fp32,
 * probabilities rounded to bf16 before P.V (A:499,512,529).  d in {64,128}. */
int lia_attn_prefill_bf16(const void* q, const void* k_cache, const void* v_cache, void* out, int B, int H, int S,
                          int d, int cache_batch, int b0, lia_stream_t stream);

/* Decode attention: one query token per sequence over T cached positions (the new token's
 * K/V row already written), NO mask (A:500).  Flash-decoding: `splits` > 1 partitions T over
 * CTAs and combines through `workspace` (lia_attn_decode_workspace_bytes()); splits == 0 lets
 * the library choose (aiming for 2 CTAs per SM), splits == -n lets it choose aiming for n CTAs
 * per SM (a tensor-parallel rank that keeps few heads).  q [B,H,d], out [B,H*d]. */
size_t lia_attn_decode_workspace_bytes(int B, int H, int d, int max_splits);
int lia_attn_decode_bf16(const void* q, const void* k_cache, const void* v_cache, void* out, int B, int H, int T,
                         int d, int cache_batch, int b0, int splits, void* workspace, size_t workspace_bytes,
                         lia_stream_t stream);

/* Stand-alone Q scaling + KV-cache append for callers that hold separate q/k/v [B,S,hq] (the
 * IndirectAccessKVCache operator face, llm/modules/mha_fusion.py:503-560); the model path fuses this
 * into the QKV GEMM epilogue.  q_out = bf16(q*q_scale); caches as in LiaQkvArgs (A:456-491). */
int lia_kv_append_bf16(const void* q, const void* k, const void* v, void* q_out, void* k_cache, void* v_cache, int B,
                       int S, int hq, int pos0, int cache_batch, int b0, float q_scale, lia_stream_t stream);

/* hidden[b,s,:] = embed_tokens[ids[b,s]] + embed_positions[past_len + s + 2]  (M:1107-1142 with
 * an all-ones attention mask, M:368-378).  ids int64 [B,S].  Either table (not both) may be NULL: its rows
 * then contribute nothing and the other table's rows are copied through -- opt-350m's project_in path looks
 * token rows [word_embed_proj_dim] and position rows [hidden] up separately (M:1139-1142). */
int lia_embed_bf16(const int64_t* ids, const void* embed_tokens, const void* embed_positions, void* out, int B,
                   int S, int h, int past_len, int vocab, int max_pos_rows, lia_stream_t stream);

/* The same with the positions OPTLearnedPositionalEmbedding.forward derives from an attention mask
 * (M:368-378): p = cumsum(mask[b,:])[past_len+s] * mask[b,past_len+s] - 1, row = p + 2.  attention_mask
 * is int64 [B, mask_ld] with mask_ld >= past_len + S (HF's mask covers past and current tokens,
 * M:1127-1131); NULL means all ones.  Padded prompts change ONLY the positions on the reference's GPU
 * branch: its attention rebuilds a pure causal mask in prefill and applies none in decode (A:446-449, A:500). */
int lia_embed_masked_bf16(const int64_t* ids, const int64_t* attention_mask, int mask_ld, const void* embed_tokens,
                          const void* embed_positions, void* out, int B, int S, int h, int past_len, int vocab,
                          int max_pos_rows, lia_stream_t stream);

/* next[b] = argmax_v logits[b, v] with logits[b, suppress_id] treated as -inf when
 * suppress_id >= 0 (min_new_tokens processor, lia/generation_utils.py:872-880; argmax at
 * intel_extension_for_pytorch/transformers/generation/greedy_search.py:395).  Lowest index
 * wins ties (torch.argmax on CPU).  logits bf16 [B,V], next int64 [B]. */
int lia_argmax_bf16(const void* logits, int64_t* next, int B, int V, int suppress_id, lia_stream_t stream);

/* out = bf16(residual + x): the residual add that follows the all-reduce of a row-parallel
 * projection under tensor parallelism (D:247, D:317; the reduce itself replaces D:60-68).
 * n = number of elements, n % 8 == 0. */
int lia_residual_add_bf16(const void* x, const void* residual, void* out, size_t n, lia_stream_t stream);

/* ---- tensor parallelism: row-parallel projection FUSED with its all-reduce and residual add.
 *
 * Replaces, for world > 1, the reference's  GPU GEMM -> .to('cpu') -> deepspeed_comm.all_reduce
 * (oneCCL) -> .to('cuda') -> residual add  chain (gpu_linear_allreduce_compute D:60-77, then
 * D:247 / D:317) with ONE kernel that exchanges partial tiles over NVLink peer memory while the
 * remaining tiles are still being computed:
 *     out = bf16( residual + bf16( sum_r  bf16( bf16(acc_r) + bias_r ) ) )
 * (bias_r is this rank's share of the bias, i.e. bias / world, tensor_parallel.py:134; the sum over
 * ranks is taken in fp32 in rank order, so every rank produces bit-identical results).
 *   M <= 128 (decode), world < 4: one-shot -- the CTA that finishes an output tile pushes it into
 *     every peer's receive area as {4 data bytes, epoch} words (data and validity arrive together:
 *     one one-way NVLink latency, no fence, no flag), then reduces the `world` partials in its own.
 *   M <= 128 (decode), world >= 4: two-shot with per-row ownership -- row group `it` of tile `ta` is
 *     reduced by rank (it + ta) % world, so every thread on every rank reduces 1/world of its own
 *     values and receives the rest as finals in the same {data, epoch} framing: two one-way
 *     latencies, 2(world-1)/world x the data per rank instead of (world-1) x.
 *   M >  128 (prefill): two-shot -- tile u is owned by rank u % world; the other ranks push their
 *     partial to the owner, which reduces, adds the residual and writes the final tile into every
 *     rank's `out` (which therefore must live inside the arena, at the same offset on every rank);
 *     flags trail their data by one tile so no warp ever waits out an NVLink round trip.  With a multicast
 *     mapping of the arena (LiaTpArgs.mc_arena) the reduction happens inside the NVLink switch instead.
 * All cross-GPU traffic goes through one symmetric "arena" per rank (lia_p2p_alloc), mapped into
 * every peer with CUDA IPC.  Every rank must issue the same sequence of calls with the same shapes.
 * Launches are safe under CUDA-graph replay (epochs live in device memory).  A peer that does
 * not show up within LIA_TP_TIMEOUT_NS makes the kernel give up and set the arena's error word
 * (lia_tp_error) instead of hanging the GPU. */
#define LIA_TP_MAX_WORLD 8
#define LIA_TP_MAX_UNITS 16384   /* output tiles per call */
typedef struct LiaTpArgs {
  int32_t rank, world;
  void* arena[LIA_TP_MAX_WORLD]; /* arena[r] = rank r's arena mapped in this process; arena[rank] is local */
  uint64_t ctl_off;              /* lia_tp_ctl_bytes() bytes, zero-initialised once (epoch, flags)      */
  uint64_t recv_off;             /* receive area: 2 * recv_bytes (two parities)                          */
  uint64_t recv_bytes;           /* >= lia_tp_recv_bytes(M, N, K, world)                                 */
  uint64_t out_off;              /* M > 128 only: offset of `out` inside the arena                       */
  void* mc_arena;                /* NVLink-switch MULTICAST address of the same arena (every rank's copy behind one
                                  * address), or NULL.  When set, the prefill exchange (M > 128) reduces INSIDE the switch:
                                  * ranks keep their partial tiles in their own arena, the owner of a tile reads the sum of
                                  * all copies with multimem.ld_reduce and writes the final tile to every rank with one
                                  * multimem.st -- ~1.1 S instead of 1.75 S bytes per rank and 7x fewer SM-issued remote
                                  * accesses.  The arena must then be symmetric memory bound to a multicast object
                                  * (torch.distributed._symmetric_memory; tp.SymmArena). */
} LiaTpArgs;
size_t lia_tp_ctl_bytes(void);
size_t lia_tp_recv_bytes(int M, int N, int K, int world);
int lia_gemm_allreduce_bf16(const void* A, const void* W, const void* bias, const void* residual, void* out, int M,
                            int N, int K, const LiaTpArgs* tp, void* workspace, size_t workspace_bytes,
                            lia_stream_t stream);
/* reads (and clears) the arena's error word: 0 = ok, 1 = a peer timed out.  Synchronises the device. */
int lia_tp_error(const LiaTpArgs* tp);

/* peer-mappable device memory (cudaMalloc + CUDA IPC).  lia_p2p_alloc zero-fills the buffer and
 * writes a 64-byte handle that another PROCESS passes to lia_p2p_open to map it. */
#define LIA_P2P_HANDLE_BYTES 64
int lia_p2p_alloc(size_t bytes, void** dev_ptr, void* handle_out);
int lia_p2p_open(const void* handle, void** peer_ptr);
int lia_p2p_close(void* peer_ptr);
int lia_p2p_free(void* dev_ptr);

/* ---- pinned host arena: replaces lia/cxl/numa_alloc.c (numa_alloc_node / numa_free_node) and
 * pin_memory (M:167-227).  Returns NULL on failure (message in lia_last_error()). */
void* lia_host_arena_alloc(size_t bytes);
int lia_host_arena_free(void* ptr, size_t bytes);

/* ---- layer streamer: double-buffered H2D of non-resident layers on a private copy stream.
 * Replaces load_layer / layer_copy (M:270-318) and the stream/event choreography of
 * M:1208-1212,1288-1316.  The caller owns the device slabs (n_slots buffers of slab_bytes). */
typedef struct LiaStreamer LiaStreamer;
LiaStreamer* lia_streamer_create(void* const* device_slabs, int n_slots, size_t slab_bytes);
/* enqueue host_src[0:bytes] -> slot on the copy stream, after all compute previously recorded
 * as using that slot (lia_streamer_release) has finished. */
int lia_streamer_prefetch(LiaStreamer* s, int slot, const void* host_src, size_t bytes);
/* make `compute_stream` wait until the last prefetch into `slot` has landed. */
int lia_streamer_wait(LiaStreamer* s, int slot, lia_stream_t compute_stream);
/* record on `compute_stream` that the slot's contents are no longer needed after this point. */
int lia_streamer_release(LiaStreamer* s, int slot, lia_stream_t compute_stream);
/* total bytes copied and device-measured copy milliseconds since creation (blocks on the copy stream). */
int lia_streamer_stats(LiaStreamer* s, double* bytes, double* copy_ms);
int lia_streamer_destroy(LiaStreamer* s);

/* ---- decode programs: a chain of decode-shaped operations (at most 128 token rows each) executed as ONE persistent
 * kernel -- what a decode step of the reference spends ~30 library launches per layer on (lia/modeling_opt.py:1379-1491
 * driving decoder.py:172-335 and attentions.py:312-557; models.py:423-431 and greedy_search.py:367-395 for the head).
 * Operations run in the order they were added; each starts when all earlier ones are complete on every SM, while the
 * weight tiles of later GEMMs and the cached K/V rows of later attention operations already stream into shared memory.
 * Every operation computes exactly what its stand-alone entry point computes (same kernels' code, same summation order):
 * results are bit-identical to calling lia_layernorm_bf16 / lia_gemm_bf16 / lia_gemm_allreduce_bf16 /
 * lia_attn_decode_bf16 (splits == 1) / lia_embed_masked_bf16 / lia_argmax_bf16 one after the other.
 * All pointers are device pointers owned by the caller and must stay valid until lia_program_destroy. */
typedef struct LiaProgram LiaProgram;
LiaProgram* lia_program_create(int rows);                       /* rows = token rows (1..128) of every operation */
int lia_program_add_layernorm(LiaProgram* p, const void* x, const void* w, const void* b, void* y, int rows, int h, float eps);
/* as lia_gemm_bf16 (tp == NULL) or lia_gemm_allreduce_bf16 (tp != NULL; `epilogue` ignored).  A LIA_EPI_QKV
 * epilogue appends at the position given to lia_program_run (qkv->pos0 is ignored). */
int lia_program_add_gemm(LiaProgram* p, const void* A, const void* W, const void* bias, const void* residual, void* out, int M,
                         int N, int K, int epilogue, const LiaQkvArgs* qkv, const LiaTpArgs* tp);
/* as lia_attn_decode_bf16 with T = pos0 + 1 of the run; cache_rows = rows the cache tensors have */
int lia_program_add_attn_decode(LiaProgram* p, const void* q, const void* k_cache, const void* v_cache, void* out, int B, int H,
                                int d, int cache_batch, int b0, int cache_rows);
/* as lia_embed_masked_bf16 with S = 1, past_len = pos0 and ids = the run's ids_in */
int lia_program_add_embed(LiaProgram* p, const int64_t* attention_mask, int mask_ld, const void* embed_tokens,
                          const void* embed_positions, void* out, int B, int h, int vocab, int max_pos_rows);
/* as lia_argmax_bf16 into the run's ids_out, with the run's suppress_id */
int lia_program_add_argmax(LiaProgram* p, const void* logits, int B, int V);
int lia_program_finalize(LiaProgram* p);                        /* uploads the program; no more operations after it */
/* one launch: pos0 = positions already cached (the new token is appended at pos0 and attends to pos0 + 1 keys) */
int lia_program_run(LiaProgram* p, int pos0, const int64_t* ids_in, int64_t* ids_out, int suppress_id, lia_stream_t stream);
int lia_program_error(LiaProgram* p);                           /* 0 = ok, 1 = a CTA timed out; synchronises the device */
int lia_program_num_ops(LiaProgram* p);
int lia_program_destroy(LiaProgram* p);

#ifdef __cplusplus
}
#endif
#endif /* LIA_B200_H_ */
