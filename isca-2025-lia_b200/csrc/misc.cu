// Small HBM-bound kernels either side of the layer stack: embeddings, greedy argmax,
// residual add after a tensor-parallel all-reduce.
#include "common.cuh"
#include "small_ops.cuh"

namespace {

// hidden[b,s,:] = embed_tokens[ids[b,s]] + embed_positions[p + 2]  (lia/modeling_opt.py:1107-1142, offset 2 at :365)
// where p is the learned-position index of OPTLearnedPositionalEmbedding.forward (lia/modeling_opt.py:368-378):
//   p = cumsum(mask[b, :])[past_len + s] * mask[b, past_len + s] - 1
// `mask` is the int64 attention mask [B, >= past_len + S] (row stride mask_ld); a null mask means all ones
// (p = past_len + s), the only case the reference's benchmark produces.
__global__ void __launch_bounds__(128) embed_kernel(const int64_t* __restrict__ ids, const bf16* __restrict__ tok,
                                                    const bf16* __restrict__ pos, bf16* __restrict__ out, int S, int h,
                                                    int past_len, int vocab, int max_pos_rows,
                                                    const int64_t* __restrict__ mask, int mask_ld) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ long long part[4];
  embed_row(ids, tok, pos, out, blockIdx.x, S, h, past_len, vocab, max_pos_rows, mask, mask_ld, threadIdx.x, part, [] { __syncthreads(); });
}

// next[b] = argmax_v logits[b,v], lowest index on ties, one id optionally suppressed
// (lia/generation_utils.py:872-880 + greedy_search.py:395)
__global__ void __launch_bounds__(256) argmax_kernel(const bf16* __restrict__ logits, int64_t* __restrict__ next, int V,
                                                     int suppress) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float sv[8];
  __shared__ int si[8];
  argmax_row<256>(logits + (size_t)blockIdx.x * V, next + blockIdx.x, V, suppress, threadIdx.x, sv, si, [] { __syncthreads(); });
}

// out = bf16(residual + x): the residual add that follows the all-reduce of a row-parallel
// projection (decoder.py:247, :317)
__global__ void __launch_bounds__(256) residual_add_kernel(const bf16* __restrict__ x, const bf16* __restrict__ res,
                                                           bf16* __restrict__ out, size_t nvec) {
  pdl_launch_dependents();
  pdl_wait();
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nvec; i += (size_t)gridDim.x * 256) {
    float a[8], c[8];
    unpack8(ldg_stream(x + i * 8), a);
    unpack8(ldg_stream(res + i * 8), c);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = c[j] + a[j];
    *reinterpret_cast<uint4*>(out + i * 8) = pack8(a);
  }
}

// q_out = bf16(q * scale); k_cache[pos0+s, b0+b] = k; v_cache[pos0+s, b0+b] = v  (attentions.py:456-491)
// for callers that computed q/k/v themselves (the IndirectAccessKVCache operator face); the model
// path fuses this into the QKV GEMM epilogue instead.
__global__ void __launch_bounds__(128) kv_append_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k,
                                                        const bf16* __restrict__ v, bf16* __restrict__ q_out,
                                                        bf16* __restrict__ kc, bf16* __restrict__ vc, int S, int hq,
                                                        int pos0, int cache_batch, int b0, float scale) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x;          // b*S + s
  const int b = row / S;
  const int s = row - b * S;
  const size_t src = (size_t)row * hq;
  const size_t dst = ((size_t)(pos0 + s) * cache_batch + b0 + b) * hq;
  for (int i = threadIdx.x * 8; i < hq; i += 128 * 8) {
    float f[8];
    unpack8(ldg_stream(q + src + i), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] *= scale;
    *reinterpret_cast<uint4*>(q_out + src + i) = pack8(f);
    *reinterpret_cast<uint4*>(kc + dst + i) = ldg_stream(k + src + i);
    *reinterpret_cast<uint4*>(vc + dst + i) = ldg_stream(v + src + i);
  }
}

}  // namespace

extern "C" int lia_kv_append_bf16(const void* q, const void* k, const void* v, void* q_out, void* k_cache, void* v_cache,
                                  int B, int S, int hq, int pos0, int cache_batch, int b0, float q_scale,
                                  lia_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LIA_CHECK_ARG(q && k && v && q_out && k_cache && v_cache, "lia_kv_append_bf16: null pointer");
  LIA_CHECK_ARG(B > 0 && S > 0 && hq > 0 && hq % 8 == 0, "lia_kv_append_bf16: bad shape");
  LIA_CHECK_ARG(pos0 >= 0 && b0 >= 0 && b0 + B <= cache_batch, "lia_kv_append_bf16: batch window outside the cache");
  lia_launch(kv_append_kernel, dim3(B * S), dim3(128), 0, stream, reinterpret_cast<const bf16*>(q), reinterpret_cast<const bf16*>(k),
                                              reinterpret_cast<const bf16*>(v), reinterpret_cast<bf16*>(q_out),
                                              reinterpret_cast<bf16*>(k_cache), reinterpret_cast<bf16*>(v_cache), S, hq, pos0,
                                              cache_batch, b0, q_scale);
  LIA_LAUNCH_CHECK();
  return LIA_OK;
}

extern "C" int lia_embed_masked_bf16(const int64_t* ids, const int64_t* attention_mask, int mask_ld, const void* embed_tokens,
                                     const void* embed_positions, void* out, int B, int S, int h, int past_len, int vocab,
                                     int max_pos_rows, lia_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LIA_CHECK_ARG(ids && out && (embed_tokens || embed_positions), "lia_embed_bf16: null pointer (at most one of the two tables may be null)");
  LIA_CHECK_ARG(B > 0 && S > 0 && h > 0 && h % 8 == 0, "lia_embed_bf16: bad shape B=%d S=%d h=%d", B, S, h);
  LIA_CHECK_ARG(past_len >= 0, "lia_embed_bf16: negative past_len");
  LIA_CHECK_ARG(embed_tokens == nullptr || vocab > 0, "lia_embed_bf16: vocab must be positive");
  LIA_CHECK_ARG(embed_positions == nullptr || past_len + S + 2 <= max_pos_rows, "lia_embed_bf16: positions %d..%d exceed the table (%d rows)", past_len + 2, past_len + S + 1, max_pos_rows);
  LIA_CHECK_ARG(attention_mask == nullptr || mask_ld >= past_len + S, "lia_embed_masked_bf16: mask rows hold %d columns, need %d", mask_ld, past_len + S);
  lia_launch(embed_kernel, dim3(B * S), dim3(128), 0, stream, ids, reinterpret_cast<const bf16*>(embed_tokens),
                                           reinterpret_cast<const bf16*>(embed_positions), reinterpret_cast<bf16*>(out), S, h,
                                           past_len, vocab, max_pos_rows, attention_mask, mask_ld);
  LIA_LAUNCH_CHECK();
  return LIA_OK;
}

extern "C" int lia_embed_bf16(const int64_t* ids, const void* embed_tokens, const void* embed_positions, void* out, int B,
                              int S, int h, int past_len, int vocab, int max_pos_rows, lia_stream_t stream_) {
  return lia_embed_masked_bf16(ids, nullptr, 0, embed_tokens, embed_positions, out, B, S, h, past_len, vocab, max_pos_rows,
                               stream_);
}

extern "C" int lia_argmax_bf16(const void* logits, int64_t* next, int B, int V, int suppress_id, lia_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LIA_CHECK_ARG(logits && next && B > 0 && V > 0, "lia_argmax_bf16: bad arguments");
  LIA_CHECK_ARG(((size_t)V * 2) % 16 == 0 || B == 1, "lia_argmax_bf16: V*2 must be a multiple of 16 bytes for B > 1 (V=%d)", V);
  lia_launch(argmax_kernel, dim3(B), dim3(256), 0, stream, reinterpret_cast<const bf16*>(logits), next, V, suppress_id);
  LIA_LAUNCH_CHECK();
  return LIA_OK;
}

extern "C" int lia_residual_add_bf16(const void* x, const void* residual, void* out, size_t n, lia_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LIA_CHECK_ARG(x && residual && out, "lia_residual_add_bf16: null pointer");
  LIA_CHECK_ARG(n % 8 == 0, "lia_residual_add_bf16: n must be a multiple of 8");
  if (n == 0) return LIA_OK;
  const size_t nvec = n / 8;
  size_t blocks = (nvec + 255) / 256;
  const size_t cap = (size_t)lia_sm_count() * 8;
  if (blocks > cap) blocks = cap;
  lia_launch(residual_add_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, reinterpret_cast<const bf16*>(x), reinterpret_cast<const bf16*>(residual),
                                                            reinterpret_cast<bf16*>(out), nvec);
  LIA_LAUNCH_CHECK();
  return LIA_OK;
}
