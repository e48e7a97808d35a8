// Internal entry points of elementwise.cu / attention.cu.
#pragma once
#include <algorithm>

#include "common.cuh"

namespace svit {

int patchify(int dtype, const float* images, void* patches, int64_t n, int C, int H, int ps, cudaStream_t stream);
int layernorm(const float* x, int64_t x_gs, int64_t x_ld, const float* gamma, const float* beta, int64_t param_gs,
              void* y, int64_t y_gs, int64_t y_ld, int out_dtype, int G, int64_t rows, int h, float eps,
              cudaStream_t stream);
int split_f16(const float* in, int64_t in_gs, void* out, int G, int64_t rows, int K, cudaStream_t stream);
int embed_cls(float* X, int64_t x_gs, const float* wvec, int64_t vec_stride, int64_t off_cls, int64_t off_pos, int G,
              int B, int T, int h, cudaStream_t stream);
int head(const float* X, int64_t x_gs, const float* wvec, int64_t vec_stride, int64_t off_g, int64_t off_b,
         int64_t off_w, int64_t off_hb, float* logits, int64_t logits_stride, int G, int B, int T, int h, int n_cls,
         float eps, cudaStream_t stream);
int attention_mma(const void* qkv, void* ctx, int dtype, int64_t n_seq, int T, int heads, cudaStream_t stream);
int attention_split(const float* qkv, float* ctx, int64_t n_seq, int T, int heads, cudaStream_t stream);
int attention_tc(const void* qkv, void* ctx, int dtype, int64_t n_seq, int T, int heads, cudaStream_t stream);
int attention_cls(const void* qkv, void* ctx_cls, int dtype, int64_t n_seq, int T, int heads, int head_dim,
                  cudaStream_t stream);
int attention(const void* qkv, void* ctx, int dtype, int64_t n_seq, int T, int heads, int head_dim,
              cudaStream_t stream);

}  // namespace svit
