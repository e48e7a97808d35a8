/*
 * svit.h -- C ABI of libsvit.so: the B200 (sm_100a) kernels behind the
 * client-contribution utility loop of shapley-vit.
 *
 * The reference (juniarto-samsudin/shapley-vit) is pure Python/PyTorch and has
 * no FFI of its own; the entry points below are what a binding for its hot path
 * replaces, function by function (paths relative to the reference root):
 *
 *   svit_aggregate        <- get_aggregated_model      federated_learning/utils.py:781-792
 *                            + ServerBase.model_agg_lazy federated_learning/server2.py:121-127
 *                            (FedAvg ratios from ServerBase.get_agg_ratio, server2.py:68-81)
 *   svit_aggregate_onto   <- the per-round accumulation of compute_utilities_lazy
 *                            fed_client_contribution/utils_fed_shapley.py:146-196 (model_agg_lazy over a list)
 *   svit_patchify,
 *   svit_forward_batched,
 *   svit_forward_lora_batched
 *                         <- net(img).logits inside evaluation, federated_learning/utils.py:886
 *                            (HF ViTForImageClassification built at start.py:258-267)
 *   svit_score, svit_score_records
 *                         <- argmax / correct / CrossEntropy(sum) in evaluation,
 *                            federated_learning/utils.py:891-894
 *   svit_plan_timing_*    <- (new) CUDA-event timing per kernel class; the reference only prints
 *                            'before net' / 'after net' (federated_learning/utils.py:885-887)
 *   svit_gemm, svit_layernorm, svit_attention
 *                         <- the ATen/cuBLAS calls the HF forward issues per layer; exported so
 *                            each kernel can be parity-tested and timed in isolation
 *   svit_layout_*         <- the state_dict key order the reference iterates
 *                            (federated_learning/utils.py:745-748, 787-791)
 *
 * Conventions: every function returns 0 on success or a negative svit_status; the message
 * of the last failure on the calling thread is svit_last_error().  No C++ exception crosses
 * the ABI.  The caller owns every buffer (device pointers, e.g. torch tensor.data_ptr());
 * the library allocates nothing on the device.  Every launch is asynchronous on the given
 * stream (a cudaStream_t, 0 = default stream) and never synchronises the host.
 * All strides are in ELEMENTS of the pointed-to type.
 */
#ifndef SVIT_H_
#define SVIT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVIT_ABI_VERSION 2

typedef void* svit_stream_t; /* cudaStream_t */

enum svit_status {
  SVIT_OK = 0,
  SVIT_ERR_ARG = -1,         /* bad argument (null pointer, size out of range) */
  SVIT_ERR_ALIGN = -2,       /* pointer / stride alignment requirement violated */
  SVIT_ERR_CUDA = -3,        /* a CUDA runtime / driver call failed */
  SVIT_ERR_UNSUPPORTED = -4, /* valid request this build cannot serve */
  SVIT_ERR_NO_DEVICE = -5    /* no sm_100 device available */
};

enum svit_dtype { SVIT_F32 = 0, SVIT_BF16 = 1, SVIT_F16 = 2 };

/* Arithmetic of the dense contractions in the forward.  Accumulation is always fp32; the
 * residual stream, LayerNorm statistics, softmax and the classifier head are always fp32. */
enum svit_precision {
  SVIT_PREC_F32 = 0,  /* fp32 operands, CUDA-core FMA GEMMs: the exact mode (small cases) */
  SVIT_PREC_TF32 = 1, /* fp32 storage, tcgen05 kind::tf32 */
  SVIT_PREC_BF16 = 2, /* bf16 operands, tcgen05 kind::f16 */
  SVIT_PREC_F16 = 3,  /* fp16 operands, tcgen05 kind::f16 (same rate as bf16, 3 more mantissa bits) */
  SVIT_PREC_F16X3 = 4, /* every GEMM operand is stored as fp16 hi + lo PLANES (SVIT_FMT_X3) and the product is
                          three tcgen05 kind::f16 passes (hi*lo + lo*hi + hi*hi) into one fp32 accumulator:
                          ~21 significant bits at a third of the f16 rate */
  SVIT_PREC_F16C8 = 5  /* fp16 main pass + fp8 COMPENSATION passes (SVIT_FMT_C8): x*y ~ hi(x)*hi(y) [kind::f16]
                          + 2^-15 (hi8(x)*lo8(y) + lo8(x)*hi8(y)) [kind::f8f6f4, e4m3, twice the f16 rate], the
                          compensation summed first and folded in by the first f16 MMA's scale-input-d:
                          ~16 significant bits (27x below one fp16 pass) for TWO f16-pass equivalents; attention
                          runs in the X3 arithmetic.  Both split modes meet the 99.9 % top-1 gate on random-init
                          weights; this one is the default */
};

/* Operand formats.  A PLAIN operand array is E elements of its svit_dtype.  The split formats are
 * PACKED OPERAND ARRAYS of several planes inside one allocation of `alloc` elements (the plane pitch;
 * alloc >= every element index used):
 *   SVIT_FMT_X3: fp16 hi = fp16(x) at byte 0, fp16 lo = fp16(x - hi) at byte 2*alloc          (4 bytes / element)
 *   SVIT_FMT_C8: fp16 hi at byte 0, e4m3 hi8 = e4m3(4 * hi) at byte 2*alloc,
 *                e4m3 lo8 = e4m3(8192 * (x - hi)) at byte 3*alloc                              (4 bytes / element)
 * Element (row, col) sits at the same index in every plane.  alloc must be a multiple of 16.
 * Range: the e4m3 planes saturate at 448, i.e. they resolve |x| <= 112; the compensation of a larger element clips
 * and its products degrade towards one fp16 pass (never below it).  fp16 `hi` itself saturates at 65 504. */
enum svit_operand_format { SVIT_FMT_PLAIN = 0, SVIT_FMT_X3 = 1, SVIT_FMT_C8 = 2 };

typedef struct svit_vit_cfg {
  int32_t hidden, layers, heads, ff, image, patch, channels, n_cls;
  float ln_eps;
} svit_vit_cfg;

/* ---- library ------------------------------------------------------------------------ */
const char* svit_version(void);
const char* svit_last_error(void);
/* Number of SMs and compute capability of the current device; SVIT_ERR_NO_DEVICE if none. */
int svit_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- plan layout -------------------------------------------------------------------- */
/* A model is one row [vec region (always fp32) | mat region (GEMM operand dtype)]; every
 * segment starts on a 64-element boundary.  Mirrors shapley_vit_b200/layout.py. */
enum svit_region { SVIT_REGION_VEC = 0, SVIT_REGION_MAT = 1 };
enum svit_seg_kind {
  SVIT_SEG_CLS = 0, SVIT_SEG_POS, SVIT_SEG_PATCH_B, SVIT_SEG_LN1_G, SVIT_SEG_LN1_B, SVIT_SEG_BQ,
  SVIT_SEG_BK, SVIT_SEG_BV, SVIT_SEG_BO, SVIT_SEG_LN2_G, SVIT_SEG_LN2_B, SVIT_SEG_B1, SVIT_SEG_B2,
  SVIT_SEG_LNF_G, SVIT_SEG_LNF_B, SVIT_SEG_HEAD_W, SVIT_SEG_HEAD_B,
  SVIT_SEG_PATCH_W, SVIT_SEG_WQ, SVIT_SEG_WK, SVIT_SEG_WV, SVIT_SEG_WO, SVIT_SEG_W1, SVIT_SEG_W2,
  SVIT_SEG_KINDS
};
typedef struct svit_segment {
  int32_t kind, layer /* -1 if not per-layer */, region;
  int32_t reserved;
  int64_t offset /* elements from the start of the region */, size, rows, cols;
} svit_segment;
int svit_layout_sizes(const svit_vit_cfg* cfg, int64_t* vec_size, int64_t* mat_size, int32_t* n_segments);
int svit_layout_segment(const svit_vit_cfg* cfg, int32_t index, svit_segment* out);

/* ---- K1: coalition aggregation -------------------------------------------------------
 * out[c, p] = cast( w0[p] + sum_{j : ratios[c,j] != 0, ascending j} ratios[c,j] * deltas[j, p] )
 * For out_dtype SVIT_F32 every product and every sum is rounded to fp32 separately (no FMA
 * contraction), i.e. exactly the arithmetic of get_aggregated_model followed by model_agg_lazy.
 * For the 16-bit out_dtypes (the operand feed of the 16-bit GEMMs; the reference has no such
 * format) the accumulation is fused and starts from w0 (a few fp32 ulp from the above before the
 * cast to 16 bits; svit_aggregate_onto keeps base + sum for every out_dtype).
 * The [N, P] stack is read from HBM ONCE for all C coalitions.
 *   deltas  device [N, delta_stride] fp32     w0   device [P] fp32 (NULL = zeros)
 *   ratios  HOST   [C, N] fp32, 0 for non-members (FedAvg n_j / sum n over the coalition); they are
 *           copied into the kernel parameters at launch, the caller may reuse the buffer at once
 *   out     device [C, out_stride] of out_dtype (SVIT_F32 / SVIT_BF16 / SVIT_F16)
 * Requirements: deltas, w0, out 16-byte aligned; delta_stride and out_stride multiples of 8
 * and >= P rounded up to 8; 1 <= N <= 64; 1 <= C <= 256. */
int svit_aggregate(const float* deltas, int64_t delta_stride, const float* w0, const float* ratios,
                   void* out, int64_t out_stride, int out_dtype, int64_t P, int N, int C,
                   svit_stream_t stream);

/* One more FL round folded onto per-coalition partial models (the reference's multi-round "lazy"
 * reconstruction: ServerBase.model_agg_lazy with a LIST of per-round aggregates, server2.py:121-127,
 * driven by compute_utilities_lazy, fed_client_contribution/utils_fed_shapley.py:146-196):
 *   out[c, p] = cast( base[c, p] + sum_{j : ratios[c,j] != 0, ascending j} ratios[c,j] * deltas[j, p] )
 * base device [C, base_stride] fp32 (may alias out when out_dtype is SVIT_F32); everything else as in
 * svit_aggregate.  A coalition without a member in this round has an all-zero ratio row: out = base. */
int svit_aggregate_onto(const float* deltas, int64_t delta_stride, const float* base, int64_t base_stride,
                        const float* ratios, void* out, int64_t out_stride, int out_dtype, int64_t P, int N,
                        int C, svit_stream_t stream);

/* The same aggregation written as a split-format packed operand array (the weight feed of SVIT_PREC_F16X3 /
 * SVIT_PREC_F16C8).  SVIT_FMT_X3 (~21 bits): the value that is split is the fp32 two-rounding result above; SVIT_FMT_C8
 * (~16 bits): the fused accumulation of the 16-bit outputs (<= 1 fp32 ulp per client from it).  w0 + sum, or
 * base[c] + sum when `base` is given -- exactly one of w0 / base may be non-NULL, both NULL = zeros.
 *   out  device packed operand array of out_alloc elements, element (c, p) at index out_off + c * out_stride + p
 *        (out_off % 8 == 0) */
int svit_aggregate_split(const float* deltas, int64_t delta_stride, const float* w0, const float* base,
                         int64_t base_stride, const float* ratios, void* out, int64_t out_off, int64_t out_stride,
                         int64_t out_alloc, int out_fmt, int64_t P, int N, int C, svit_stream_t stream);

/* ---- forward ------------------------------------------------------------------------ */
typedef struct svit_plan svit_plan; /* opaque: geometry, layout table, TMA descriptors */
int svit_plan_create(const svit_vit_cfg* cfg, int precision, int max_coalitions, int max_images,
                     svit_plan** out);
int svit_plan_destroy(svit_plan* plan);
/* Bytes of caller-provided device workspace svit_forward_batched needs for (C, B) up to the
 * plan's maxima; 256-byte aligned. */
int64_t svit_plan_workspace_bytes(const svit_plan* plan);
/* dtype (svit_dtype) and format (svit_operand_format) of the mat region / patch matrix for this plan's
 * precision (split formats: dtype is SVIT_F16, 4 bytes per element over all planes). */
int svit_plan_operand_dtype(const svit_plan* plan);
int svit_plan_operand_format(const svit_plan* plan);

/* images [n, channels, image, image] fp32 (NCHW) -> rows [row0, row0 + n * n_patches) of the patch matrix
 * [rows, channels*patch*patch] in the plan's operand format (split formats: a packed operand array of
 * patches_alloc elements), column order (channel, row, col) = the flattened conv kernel. */
int svit_patchify(const svit_plan* plan, const float* images, void* patches, int64_t patches_alloc, int64_t row0,
                  int64_t n, svit_stream_t stream);

/* logits[c * logits_stride + b * n_cls + k] for c < C, b < B: the ViT forward of image b under
 * the weights of coalition c.
 *   wvec  device [C, vec_stride] fp32            (vec region, from svit_aggregate)
 *   wmat  device [C, mat_stride] operand format  (mat region, from svit_aggregate / svit_aggregate_split;
 *         split formats: packed operand array of wmat_alloc elements)
 *   patches device [.., patch_dim] operand format (from svit_patchify; packed array of patches_alloc elements),
 *         the B images start at row patches_row0; shared by all C */
int svit_forward_batched(svit_plan* plan, const float* wvec, int64_t vec_stride, const void* wmat,
                         int64_t mat_stride, int64_t wmat_alloc, const void* patches, int64_t patches_alloc,
                         int64_t patches_row0, float* logits, int64_t logits_stride, int C, int B, void* workspace,
                         size_t workspace_bytes, svit_stream_t stream);

/* The same forward for PEFT-LoRA coalition models over a FROZEN base (reference start.py:274-283: LoRA r = 16 on query /
 * value, classifier in modules_to_save): every weight matrix is shared by all C coalitions -- ONE mat region, every
 * GEMM takes one B operand for all groups -- and the wrapped projections add the per-coalition low-rank term
 *     q = x Wq^T + bq + (alpha / r) (x A_q^T) B_q^T          (PEFT lora.Linear, eval mode; likewise v)
 * as a K-extension of the shared QKV GEMM.  FedAvg averages A and B separately (they are state_dict entries), so the
 * term is not linear in the coalition: the caller aggregates the factors (svit_aggregate / svit_aggregate_split on the
 * packed LoRA rows) into
 *   lora  device [C, lora_stride] in the plan's operand format (packed array of lora_alloc elements), per layer l at
 *         element l * 256 * hidden:  Acat [64, hidden] (rows 0 .. r-1 = A_q, rows 32 .. 32+r-1 = A_v, others 0), then
 *         Bext [3 * hidden, 64] (query rows: columns 0 .. r-1 = (alpha/r) B_q; value rows: columns 32 .. = (alpha/r) B_v;
 *         key rows 0);  r <= 32
 *   wmat_shared  device [mat_size] in the operand format (one row);  wvec as in svit_forward_batched (per coalition:
 *         the classifier and any bias / LayerNorm parameter the clients did train). */
int svit_forward_lora_batched(svit_plan* plan, const float* wvec, int64_t vec_stride, const void* wmat_shared,
                              int64_t wmat_alloc, const void* lora, int64_t lora_stride, int64_t lora_alloc,
                              const void* patches, int64_t patches_alloc, int64_t patches_row0, float* logits,
                              int64_t logits_stride, int C, int B, void* workspace, size_t workspace_bytes,
                              svit_stream_t stream);

/* Device-side timing of the forward by kernel class, for roofline reporting: between _begin and
 * _end every launch of svit_forward_batched on this plan is bracketed by CUDA events on its
 * stream; _end waits for them and returns, per class, the summed duration, the summed
 * algorithmic work (FLOPs for GEMM and attention, bytes for LayerNorm) and the launch count.
 * SVIT_CLS_FORWARD spans whole svit_forward_batched calls. */
enum svit_kernel_class { SVIT_CLS_GEMM = 0, SVIT_CLS_ATTENTION, SVIT_CLS_LAYERNORM, SVIT_CLS_FORWARD, SVIT_CLS_COUNT };
typedef struct svit_timing {
  double ms[SVIT_CLS_COUNT];
  double work[SVIT_CLS_COUNT];
  int64_t launches[SVIT_CLS_COUNT];
} svit_timing;
int svit_plan_timing_begin(svit_plan* plan);
int svit_plan_timing_end(svit_plan* plan, svit_timing* out);

/* ---- K5: scoring ---------------------------------------------------------------------
 * For each coalition c: correct[c] (+)= #{i : argmax_k logits[c,i,k] == labels[i]} (first maximal
 * index on ties, like torch.argmax), loss_sum[c] (+)= sum_i CE(logits[c,i,:], labels[i]) with fp32
 * log-sum-exp per sample and a deterministic fp64 reduction.  pred (optional) receives the argmax.
 *   logits device [C, logits_stride] fp32, rows of n_cls     labels device [n] int64
 *   accumulate != 0 adds to correct / loss_sum instead of overwriting. */
int svit_score(const float* logits, int64_t logits_stride, const int64_t* labels, int C, int64_t n,
               int n_cls, int64_t* correct, double* loss_sum, int32_t* pred, int64_t pred_stride,
               int accumulate, svit_stream_t stream);

/* The same scoring written as packed 16-byte records, the element of the multi-GPU exchange: a rank's K5 launches
 * write its coalitions' records straight into its slot of the all-gather send buffer (one
 * ncclAllGather of per-coalition (correct:int64, loss_sum:fp64) pairs per wave; no host round trip in between).
 *   records device [C] svit_record, 16-byte aligned */
typedef struct svit_record {
  int64_t correct;
  double loss_sum;
} svit_record;
int svit_score_records(const float* logits, int64_t logits_stride, const int64_t* labels, int C, int64_t n,
                       int n_cls, svit_record* records, int accumulate, svit_stream_t stream);

/* ---- building blocks (exported for per-kernel parity tests and roofline timing) ------ */
typedef struct svit_epilogue {
  const float* bias;     int64_t bias_gs;     /* [N] per group, NULL = none */
  const float* rowvec;   int64_t rowvec_gs;   /* [rows_out, N] per group, added by (out_row % rows_out); NULL = none */
  const float* residual; int64_t residual_gs; /* fp32 [M_out, N] per group, added last; may alias out */
  int32_t gelu;                               /* exact-erf GELU after bias */
  int32_t rows_in, rows_out, row_shift;       /* rows_in > 0: out_row = (r / rows_in) * rows_out + row_shift + r % rows_in */
} svit_epilogue;

/* For g < G: out[g] = epilogue( A[g] (M x K, row-major) * B[g]^T (B[g] is N x K, row-major) ).
 * A, B in the operand dtype of `precision`; out_dtype is SVIT_F32 or that operand dtype.
 * Group strides may be 0 (operand shared by all groups: A and, on the tensor-core precisions, B).  M_out = M unless rows_in > 0. */
/* SVIT_PREC_F16X3 / SVIT_PREC_F16C8: A and B are packed operand arrays (svit_split_operand) of exactly
 * (a_gs ? G : 1) * M * K and G * N * K elements in the precision's format; out_dtype SVIT_F32 gives a plain fp32
 * result, out_dtype SVIT_F16 a packed operand array of G * M_out * N elements in the same format (no rowvec /
 * residual then). */
int svit_gemm(int precision, const void* A, int64_t a_gs, const void* B, int64_t b_gs, void* out,
              int64_t out_gs, int out_dtype, int G, int M, int N, int K, const svit_epilogue* epi,
              svit_stream_t stream);

/* out[g] = epilogue( A[g] B[g]^T + Ae[g] Be[g]^T ): svit_gemm with a K-EXTENSION of 64 columns -- Ae [G, M, 64] and
 * Be [G, N, 64] in the operand format of `precision` (a tensor-core precision) -- accumulated in the same pass as
 * extra k-blocks.  b_gs may be 0 (B shared by all groups): the per-group low-rank correction of a shared weight. */
int svit_gemm_ext(int precision, const void* A, int64_t a_gs, const void* B, int64_t b_gs, const void* Ae,
                  const void* Be, void* out, int64_t out_gs, int out_dtype, int G, int M, int N, int K,
                  const svit_epilogue* epi, svit_stream_t stream);

/* fp32 array of `elems` elements -> packed operand array (out_fmt = SVIT_FMT_X3 / SVIT_FMT_C8) of out_alloc >=
 * elems elements: what the producers of the forward (svit_aggregate_split, LayerNorm, the GEMM epilogues,
 * attention) emit directly.  elems % 4 == 0, out_alloc % 16 == 0, 16-byte aligned pointers. */
int svit_split_operand(const float* in, void* out, int64_t out_alloc, int out_fmt, int64_t elems,
                       svit_stream_t stream);

/* y[g, r, :] = LayerNorm(x[g, r, :]) * gamma[g] + beta[g]; x fp32 [G, rows, h] (row stride x_ld),
 * y in out_dtype (out_fmt SVIT_FMT_PLAIN) or a packed operand array of out_alloc elements (split formats),
 * biased variance, eps as given (HF ViT: 1e-12). */
int svit_layernorm(const float* x, int64_t x_gs, int64_t x_ld, const float* gamma, const float* beta,
                   int64_t param_gs, void* y, int64_t y_gs, int64_t y_ld, int out_dtype, int out_fmt,
                   int64_t out_alloc, int G, int64_t rows, int h, float eps, svit_stream_t stream);

/* qkv [n_seq, T, 3h] (q | k | v, heads contiguous inside each) in `dtype` ->
 * ctx [n_seq, T, h] = softmax(q k^T / sqrt(d)) v per head, fp32 softmax. */
int svit_attention(const void* qkv, void* ctx, int dtype, int64_t n_seq, int T, int heads, int head_dim,
                   svit_stream_t stream);
/* The attention of the split precisions: qkv is an SVIT_FMT_X3 packed operand array of n_seq * T * 3h elements,
 * ctx a packed operand array (ctx_fmt = SVIT_FMT_X3 / SVIT_FMT_C8) of n_seq * T * h elements.  Every product runs
 * as fp16 hi*lo + lo*hi + hi*hi with fp32 accumulation (tcgen05 for head_dim 64 and 128 < T <= 216, CUDA cores
 * otherwise), fp32 softmax. */
int svit_attention_split(const void* qkv, void* ctx, int ctx_fmt, int64_t n_seq, int T, int heads, int head_dim,
                         svit_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SVIT_H_ */
