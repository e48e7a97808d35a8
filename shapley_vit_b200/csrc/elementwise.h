// Internal entry points of elementwise.cu / attention.cu.
#pragma once
#include <algorithm>

#include "common.cuh"

namespace svit {

// outputs: `Operand` = a plain array of out_dtype or a split-format packed operand array (include/svit.h)
int patchify_at(int out_dtype, const float* images, const Operand& patches, int64_t off, int64_t n, int C, int H, int ps,
                cudaStream_t stream);
int layernorm(const float* x, int64_t x_gs, int64_t x_ld, const float* gamma, const float* beta, int64_t param_gs,
              const Operand& y, int64_t y_gs, int64_t y_ld, int out_dtype, int G, int64_t rows, int h, float eps,
              cudaStream_t stream);
int split_operand(const float* in, const Operand& out, int64_t elems, cudaStream_t stream);
int embed_cls(float* X, int64_t x_gs, const float* wvec, int64_t vec_stride, int64_t off_cls, int64_t off_pos, int G,
              int B, int T, int h, cudaStream_t stream);
int head(const float* X, int64_t x_gs, const float* wvec, int64_t vec_stride, int64_t off_g, int64_t off_b,
         int64_t off_w, int64_t off_hb, float* logits, int64_t logits_stride, int G, int B, int T, int h, int n_cls,
         float eps, cudaStream_t stream);
int attention_mma(const void* qkv, void* ctx, int dtype, int64_t n_seq, int T, int heads, cudaStream_t stream);
// split-precision attention: qkv is an X3 packed operand array [n_seq, T, 3h], ctx a split-format array [n_seq, T, h]
// (cls_only: ctx [n_seq, h], the [CLS] query alone); every product carries ~21 bits, fp32 softmax
int attention_split(const Operand& qkv, const Operand& ctx, int64_t n_seq, int T, int heads, int head_dim, bool cls_only,
                    cudaStream_t stream);
int attention_split_tc(const Operand& qkv, const Operand& ctx, int64_t n_seq, int T, int heads, cudaStream_t stream);
bool attention_split_tc_fits(int T);
int attention_tc(const void* qkv, void* ctx, int dtype, int64_t n_seq, int T, int heads, cudaStream_t stream);
int attention_cls(const void* qkv, void* ctx_cls, int dtype, int64_t n_seq, int T, int heads, int head_dim,
                  cudaStream_t stream);
int attention(const void* qkv, void* ctx, int dtype, int64_t n_seq, int T, int heads, int head_dim,
              cudaStream_t stream);

}  // namespace svit
