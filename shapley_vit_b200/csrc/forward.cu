// Batched multi-coalition ViT forward: svit_plan_*, svit_patchify, svit_forward_batched,
// svit_gemm.
//
// Replaces `net(img).logits` inside the reference's evaluation loop
// (federated_learning/utils.py:880-886) for C coalition models at once.  Every dense layer is a
// GROUPED GEMM over the coalition axis (group g reads the g-th row of the aggregated weight
// matrix region written by svit_aggregate); activations of all groups live side by side in one
// caller-provided workspace:
//
//   X    fp32     [C, B*T, h]   residual stream (always fp32)
//   Xn   operand  [C, B*T, h]   LayerNorm output (GEMM A operand)
//   QKV  operand  [C, B*T, 3h]
//   CTX  operand  [C, B*T, h]   attention output
//   H    operand  [C, B*T, ff]  GELU(MLP up)
//
// "operand" = the plan's operand format: a plain fp32 / bf16 / fp16 array, or -- in the split precisions
// SVIT_PREC_F16X3 / F16C8 -- a packed operand array of planes (include/svit.h) that the PRODUCER of each
// tensor writes directly (svit_aggregate_split, LayerNorm, the GEMM epilogues, attention): no separate
// split pass, no fp32 copy of any activation.  QKV is always X3 planes there (attention runs hi / lo fp16
// in both split precisions); Xn, CTX and H are in the GEMMs' format.
//
// Per layer: LN -> QKV GEMM(+bias) -> fused-softmax attention -> proj GEMM(+bias +residual, in
// place on X) -> LN -> MLP-up GEMM(+bias +GELU) -> MLP-down GEMM(+bias +residual).  The patch
// embedding is one GEMM whose A operand (the patchified validation images) is shared by all
// coalitions; its epilogue adds the conv bias and the position embedding and leaves the [CLS]
// row gap.  The head (final LN on CLS + classifier) is one fp32 warp per (coalition, image).
#include <cstdlib>
#include <new>
#include <vector>

#include "elementwise.h"
#include "epilogue.cuh"
#include "layout.h"

struct svit_plan {
  svit::Layout lay;
  int precision = 0, operand_dtype = 0;
  int max_c = 0, max_b = 0;
  int T = 0, np = 0, pd = 0;
  bool force_simt = false;
  bool full_last_layer = false;  // SVIT_FULL_LAST_LAYER=1: compute every token of the last layer (A/B and tests)
  // workspace byte offsets for (max_c, max_b)
  size_t off_x = 0, off_xn = 0, off_qkv = 0, off_ctx = 0, off_h = 0, off_t = 0, ws_bytes = 0;
  int fmt = 0;                    // svit_operand_format of the GEMM operands (weights, patches, Xn, CTX, H)
  int64_t rows_max = 0;           // max_c * max_b * T: the plane pitches of the activation arrays are rows_max * width
  // optional per-kernel-class device timing (svit_plan_timing_begin / _end)
  bool timing = false;
  struct Span {
    cudaEvent_t a, b;
    int cls;
    double work;  // flops (GEMM, attention) or bytes (others)
  };
  std::vector<Span> spans;
  std::vector<cudaEvent_t> pool;
  size_t pool_used = 0;
};

namespace svit {
namespace {

int operand_dtype_of(int precision) {
  switch (precision) {
    case SVIT_PREC_F32:
    case SVIT_PREC_TF32: return SVIT_F32;
    case SVIT_PREC_BF16: return SVIT_BF16;
    case SVIT_PREC_F16:
    case SVIT_PREC_F16X3:
    case SVIT_PREC_F16C8: return SVIT_F16;  // (split precisions: the main plane)
    default: return -1;
  }
}

cudaEvent_t take_event(svit_plan* p) {
  if (p->pool_used == p->pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    p->pool.push_back(e);
  }
  return p->pool[p->pool_used++];
}

// RAII span: records an event pair around the launches issued in its scope when timing is on
struct Timed {
  svit_plan* p;
  cudaStream_t s;
  svit_plan::Span span;
  Timed(svit_plan* plan, cudaStream_t stream, int cls, double work) : p(plan), s(stream) {
    if (!p->timing) return;
    span.a = take_event(p);
    span.b = take_event(p);
    span.cls = cls;
    span.work = work;
    cudaEventRecord(span.a, s);
  }
  ~Timed() {
    if (!p->timing) return;
    cudaEventRecord(span.b, s);
    p->spans.push_back(span);
  }
};

int gemm_dispatch(svit_plan* p, const Operand& A, int64_t a_off, int64_t a_gs, const Operand& B, int64_t b_off, int64_t b_gs,
                  int G, int M, int N, int K, const EpiArgs& epi, cudaStream_t stream, const GemmExt* ext = nullptr) {
  Timed t(p, stream, SVIT_CLS_GEMM, 2.0 * G * M * (double)N * (K + (ext ? kGemmExtK : 0)));
  if (p->precision == SVIT_PREC_F32 || (p->force_simt && p->fmt == SVIT_FMT_PLAIN)) {
    const int es = dtype_size(p->operand_dtype);
    int rc = gemm_simt(p->operand_dtype, A.plane(0, a_off, es), a_gs, B.plane(0, b_off, es), b_gs, G, M, N, K, epi, stream);
    if (rc || !ext) return rc;
    // CUDA-core path: the K-extension as a second GEMM accumulating in place (plain output arrays only)
    if (epi.out_dtype != p->operand_dtype && epi.out_dtype != SVIT_F32) SVIT_FAIL(SVIT_ERR_UNSUPPORTED, "gemm: extension output dtype");
    EpiArgs e2{};
    e2.out = epi.out, e2.out_gs = epi.out_gs, e2.out_dtype = epi.out_dtype, e2.M = epi.M, e2.N = epi.N;
    e2.residual_gs = epi.out_gs;
    if (epi.out_dtype != SVIT_F32) SVIT_FAIL(SVIT_ERR_UNSUPPORTED, "gemm: the CUDA-core K-extension needs an fp32 output");
    e2.residual = static_cast<const float*>(epi.out);
    return gemm_simt(p->operand_dtype, ext->A.plane(0, ext->a_off, es), ext->a_gs, ext->B.plane(0, ext->b_off, es), ext->b_gs, G, M,
                     N, kGemmExtK, e2, stream);
  }
  return gemm_tc(p->precision, A, a_off, a_gs, B, b_off, b_gs, G, M, N, K, epi, stream, ext);
}

}  // namespace
}  // namespace svit

extern "C" int svit_plan_create(const svit_vit_cfg* cfg, int precision, int max_coalitions, int max_images,
                                svit_plan** out) {
  using namespace svit;
  SVIT_CHECK_ARG(out != nullptr, "svit_plan_create: out is null");
  *out = nullptr;
  SVIT_CHECK_ARG(max_coalitions >= 1 && max_images >= 1, "svit_plan_create: max_coalitions/max_images must be >= 1");
  const int odt = operand_dtype_of(precision);
  SVIT_CHECK_ARG(odt >= 0, "svit_plan_create: unknown precision %d", precision);
  svit_plan* p = new (std::nothrow) svit_plan();
  SVIT_CHECK_ARG(p != nullptr, "svit_plan_create: out of host memory");
  int rc = build_layout(cfg, &p->lay);
  if (rc) {
    delete p;
    return rc;
  }
  if (cfg->hidden > 1024 || (cfg->hidden / cfg->heads) % 32 != 0) {
    delete p;
    SVIT_FAIL(SVIT_ERR_UNSUPPORTED, "svit_plan_create: hidden <= 1024 and head_dim %% 32 == 0 required");
  }
  p->precision = precision;
  p->operand_dtype = odt;
  p->fmt = format_of_precision(precision);
  if (p->fmt != SVIT_FMT_PLAIN && cfg->hidden / cfg->heads != 64) {
    delete p;
    SVIT_FAIL(SVIT_ERR_UNSUPPORTED, "svit_plan_create: the split precisions take head_dim 64");
  }
  p->max_c = max_coalitions;
  p->max_b = max_images;
  p->np = (cfg->image / cfg->patch) * (cfg->image / cfg->patch);
  p->T = p->np + 1;
  p->pd = cfg->channels * cfg->patch * cfg->patch;
  const char* env = getenv("SVIT_FORCE_SIMT_GEMM");  // debugging aid: CUDA-core GEMMs on 16-bit operands
  p->force_simt = env && env[0] == '1';
  const char* fl = getenv("SVIT_FULL_LAST_LAYER");
  p->full_last_layer = fl && fl[0] == '1';
  const size_t rows = (size_t)max_coalitions * max_images * p->T;
  p->rows_max = (int64_t)rows;
  const size_t es = p->fmt == SVIT_FMT_PLAIN ? (size_t)dtype_size(odt) : 4;  // bytes per element over all planes
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    return o;
  };
  p->off_x = take(rows * cfg->hidden * 4);
  p->off_xn = take(rows * cfg->hidden * es);
  p->off_qkv = take(rows * 3 * cfg->hidden * es);
  p->off_ctx = take(rows * cfg->hidden * es);
  p->off_h = take(rows * cfg->ff * es);
  p->off_t = take(rows * kGemmExtK * es);  // LoRA path: T = Xn [A_q; A_v]^T, the A operand of the QKV K-extension
  p->ws_bytes = off;
  *out = p;
  return SVIT_OK;
}

extern "C" int svit_plan_destroy(svit_plan* plan) {
  if (plan)
    for (cudaEvent_t e : plan->pool) cudaEventDestroy(e);
  delete plan;
  return SVIT_OK;
}

extern "C" int64_t svit_plan_workspace_bytes(const svit_plan* plan) { return plan ? (int64_t)plan->ws_bytes : -1; }
extern "C" int svit_plan_operand_dtype(const svit_plan* plan) { return plan ? plan->operand_dtype : -1; }
extern "C" int svit_plan_operand_format(const svit_plan* plan) { return plan ? plan->fmt : -1; }

extern "C" int svit_patchify(const svit_plan* plan, const float* images, void* patches, int64_t patches_alloc, int64_t row0,
                             int64_t n, svit_stream_t stream) {
  using namespace svit;
  SVIT_CHECK_ARG(plan && images && patches && n >= 0 && row0 >= 0, "svit_patchify: bad arguments");
  if (!aligned16(images) || !aligned16(patches)) SVIT_FAIL(SVIT_ERR_ALIGN, "svit_patchify: pointers must be 16-byte aligned");
  const svit_vit_cfg& c = plan->lay.cfg;
  const int64_t pd = plan->pd;
  SVIT_CHECK_ARG(plan->fmt == SVIT_FMT_PLAIN || (patches_alloc % 16 == 0 && patches_alloc >= (row0 + n * plan->np) * pd),
                 "svit_patchify: plane pitch %lld does not cover the rows written", (long long)patches_alloc);
  Operand dst{patches, plan->fmt, patches_alloc};
  int64_t off = row0 * pd;
  if (plan->fmt == SVIT_FMT_PLAIN) {  // plain array: just a pointer
    dst.base = dst.plane(0, off, dtype_size(plan->operand_dtype));
    off = 0;
  }
  return patchify_at(plan->operand_dtype, images, dst, off, n, c.channels, c.image, c.patch, static_cast<cudaStream_t>(stream));
}

namespace svit {
namespace {
// per-coalition LoRA operand rows (svit_forward_lora_batched): layer l at element l * (64 h + 3h 64) of a row:
// Acat [64, h] then Bext [3h, 64], both K-major, in the plan's operand format
struct LoraArgs {
  const void* rows;
  int64_t stride, alloc;
};
int forward_impl(svit_plan* plan, const float* wvec, int64_t vec_stride, const void* wmat, int64_t mat_stride,
                 int64_t wmat_alloc, const void* patches, int64_t patches_alloc, int64_t patches_row0, float* logits,
                 int64_t logits_stride, int C, int B, void* workspace, size_t workspace_bytes, svit_stream_t stream_,
                 const LoraArgs* lora);
}  // namespace
}  // namespace svit

extern "C" int svit_forward_batched(svit_plan* plan, const float* wvec, int64_t vec_stride, const void* wmat,
                                    int64_t mat_stride, int64_t wmat_alloc, const void* patches, int64_t patches_alloc,
                                    int64_t patches_row0, float* logits, int64_t logits_stride, int C, int B,
                                    void* workspace, size_t workspace_bytes, svit_stream_t stream_) {
  return svit::forward_impl(plan, wvec, vec_stride, wmat, mat_stride, wmat_alloc, patches, patches_alloc, patches_row0, logits,
                            logits_stride, C, B, workspace, workspace_bytes, stream_, nullptr);
}

extern "C" int svit_forward_lora_batched(svit_plan* plan, const float* wvec, int64_t vec_stride, const void* wmat_shared,
                                         int64_t wmat_alloc, const void* lora, int64_t lora_stride, int64_t lora_alloc,
                                         const void* patches, int64_t patches_alloc, int64_t patches_row0, float* logits,
                                         int64_t logits_stride, int C, int B, void* workspace, size_t workspace_bytes,
                                         svit_stream_t stream_) {
  using namespace svit;
  SVIT_CHECK_ARG(plan && lora, "svit_forward_lora_batched: null pointer");
  const svit_vit_cfg& cfg = plan->lay.cfg;
  const int64_t per_layer = (int64_t)kGemmExtK * cfg.hidden * 4;  // 64 h (Acat) + 3h 64 (Bext)
  SVIT_CHECK_ARG(lora_stride >= per_layer * cfg.layers && lora_stride % 64 == 0,
                 "svit_forward_lora_batched: lora_stride must cover layers * 256 * hidden elements and be a multiple of 64");
  SVIT_CHECK_ARG(plan->fmt == SVIT_FMT_PLAIN || (lora_alloc % 16 == 0 && lora_alloc >= (int64_t)(C - 1) * lora_stride + per_layer * cfg.layers),
                 "svit_forward_lora_batched: bad plane pitch for the LoRA rows");
  if (!aligned16(lora)) SVIT_FAIL(SVIT_ERR_ALIGN, "svit_forward_lora_batched: lora must be 16-byte aligned");
  LoraArgs la{lora, lora_stride, lora_alloc};
  return forward_impl(plan, wvec, vec_stride, wmat_shared, 0, wmat_alloc, patches, patches_alloc, patches_row0, logits,
                      logits_stride, C, B, workspace, workspace_bytes, stream_, &la);
}

namespace svit {
namespace {
int forward_impl(svit_plan* plan, const float* wvec, int64_t vec_stride, const void* wmat, int64_t mat_stride,
                 int64_t wmat_alloc, const void* patches, int64_t patches_alloc, int64_t patches_row0, float* logits,
                 int64_t logits_stride, int C, int B, void* workspace, size_t workspace_bytes, svit_stream_t stream_,
                 const LoraArgs* lora) {
  SVIT_CHECK_ARG(plan && wvec && wmat && patches && logits && workspace, "svit_forward_batched: null pointer");
  SVIT_CHECK_ARG(C >= 1 && C <= plan->max_c && B >= 1 && B <= plan->max_b,
                 "svit_forward_batched: (C=%d, B=%d) exceeds the plan's (%d, %d)", C, B, plan->max_c, plan->max_b);
  SVIT_CHECK_ARG(workspace_bytes >= plan->ws_bytes, "svit_forward_batched: workspace too small (%zu < %zu)",
                 workspace_bytes, plan->ws_bytes);
  const svit_vit_cfg& cfg = plan->lay.cfg;
  const Layout& L = plan->lay;
  // mat_stride == 0 (LoRA path only): ONE weight matrix region shared by every coalition
  SVIT_CHECK_ARG(vec_stride >= L.vec_size && (mat_stride >= L.mat_size || (lora && mat_stride == 0)) && vec_stride % 64 == 0 &&
                     mat_stride % 64 == 0,
                 "svit_forward_batched: weight strides must be multiples of 64 and cover the layout");
  SVIT_CHECK_ARG(logits_stride >= (int64_t)B * cfg.n_cls, "svit_forward_batched: logits_stride too small");
  if (!aligned16(wvec) || !aligned16(wmat) || !aligned16(patches) || ((uintptr_t)workspace & 255))
    SVIT_FAIL(SVIT_ERR_ALIGN, "svit_forward_batched: weights/patches must be 16-byte and workspace 256-byte aligned");
  const int fmt = plan->fmt;
  const bool split = fmt != SVIT_FMT_PLAIN;
  SVIT_CHECK_ARG(!split || (wmat_alloc % 16 == 0 && patches_alloc % 16 == 0 && wmat_alloc >= (int64_t)(mat_stride ? C - 1 : 0) * mat_stride + L.mat_size),
                 "svit_forward_batched: bad plane pitches for the split-format weights / patches");

  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int h = cfg.hidden, ff = cfg.ff, T = plan->T, np = plan->np, pd = plan->pd;
  const int M = B * T;
  const int odt = plan->operand_dtype;
  char* ws = static_cast<char*>(workspace);
  float* X = reinterpret_cast<float*>(ws + plan->off_x);
  // activation arrays: their plane pitch is the plan's maximum row count times the width
  const Operand Xn{ws + plan->off_xn, fmt, plan->rows_max * h};
  const Operand QKV{ws + plan->off_qkv, split ? SVIT_FMT_X3 : SVIT_FMT_PLAIN, plan->rows_max * 3 * h};
  const Operand CTX{ws + plan->off_ctx, fmt, plan->rows_max * h};
  const Operand H{ws + plan->off_h, fmt, plan->rows_max * ff};
  const Operand TT{ws + plan->off_t, fmt, plan->rows_max * kGemmExtK};  // [C, M, 64] (LoRA path)
  const Operand LR{const_cast<void*>(lora ? lora->rows : nullptr), fmt, lora ? lora->alloc : 0};
  const Operand WM{const_cast<void*>(wmat), fmt, wmat_alloc};
  const Operand PA{const_cast<void*>(patches), fmt, patches_alloc};
  const int es = dtype_size(odt);
  auto mat = [&](int kind, int layer) -> int64_t { return L.find(kind, layer); };  // element offset inside a coalition's mat row
  auto vec = [&](int kind, int layer) -> const float* { return wvec + L.find(kind, layer); };
  auto out_op = [&](EpiArgs& ea, const Operand& o) { if (split) set_out_format(ea, o.fmt, o.alloc); };
  const int64_t xgs = (int64_t)M * h;
  int rc;

  // ---- embeddings ----
  Timed whole(plan, stream, SVIT_CLS_FORWARD, 0.0);
  if ((rc = embed_cls(X, xgs, wvec, vec_stride, L.find(SVIT_SEG_CLS), L.find(SVIT_SEG_POS), C, B, T, h, stream))) return rc;
  {
    svit_epilogue e{};
    e.bias = vec(SVIT_SEG_PATCH_B, -1);
    e.bias_gs = vec_stride;
    e.rowvec = vec(SVIT_SEG_POS, -1);
    e.rowvec_gs = vec_stride;
    e.rows_in = np;
    e.rows_out = T;
    e.row_shift = 1;
    EpiArgs ea = make_epi(&e, X, xgs, SVIT_F32, B * np, h);
    if ((rc = gemm_dispatch(plan, PA, patches_row0 * pd, 0, WM, mat(SVIT_SEG_PATCH_W, -1), mat_stride, C, B * np, h, pd, ea, stream)))
      return rc;
  }
  // ---- encoder ----
  for (int l = 0; l < cfg.layers; ++l) {
    {
      Timed t(plan, stream, SVIT_CLS_LAYERNORM, (double)C * M * h * (4.0 + (split ? 4 : es)));
      if ((rc = layernorm(X, xgs, h, vec(SVIT_SEG_LN1_G, l), vec(SVIT_SEG_LN1_B, l), vec_stride, Xn, xgs, h, odt, C, M, h,
                          cfg.ln_eps, stream)))
        return rc;
    }
    {
      svit_epilogue e{};
      e.bias = vec(SVIT_SEG_BQ, l);  // bq | bk | bv are contiguous
      e.bias_gs = vec_stride;
      EpiArgs ea = make_epi(&e, QKV.base, (int64_t)M * 3 * h, odt, M, 3 * h);
      out_op(ea, QKV);
      if (!lora) {
        if ((rc = gemm_dispatch(plan, Xn, 0, xgs, WM, mat(SVIT_SEG_WQ, l), mat_stride, C, M, 3 * h, h, ea, stream))) return rc;
      } else {
        // PEFT lora.Linear on query / value (reference start.py:274-276): q = x Wq^T + bq + (alpha/r) (x A_q^T) B_q^T.
        // T = Xn Acat^T (rank-64 block: rows 0.. of A_q, 32.. of A_v), then the shared-weight QKV GEMM with the
        // K-extension T Bext^T (Bext rows of q / v hold the scaled B factors, the k rows are zero).
        const int64_t lo = (int64_t)l * kGemmExtK * h * 4, tgs = (int64_t)M * kGemmExtK;
        EpiArgs et = make_epi(nullptr, TT.base, tgs, odt, M, kGemmExtK);
        out_op(et, TT);
        if ((rc = gemm_dispatch(plan, Xn, 0, xgs, LR, lo, lora->stride, C, M, kGemmExtK, h, et, stream))) return rc;
        const GemmExt ext{TT, 0, tgs, LR, lo + (int64_t)kGemmExtK * h, lora->stride};
        if ((rc = gemm_dispatch(plan, Xn, 0, xgs, WM, mat(SVIT_SEG_WQ, l), 0, C, M, 3 * h, h, ea, stream, &ext))) return rc;
      }
    }
    if (l == cfg.layers - 1 && !plan->full_last_layer) {
      // Last layer: the classifier reads only the [CLS] token (HF modeling_vit.py:641-642), so past
      // K and V only the [CLS] rows are computed: attention of the [CLS] query, then out-proj, LN,
      // MLP on B rows per coalition instead of B * T.  Same operations on those rows, same logits.
      const int64_t cgs = (int64_t)B * h;  // compact [C, B, h] buffers live at the front of CTX / Xn / H
      {
        Timed t(plan, stream, SVIT_CLS_ATTENTION, 4.0 * C * B * (double)T * h);
        if (split) rc = attention_split(QKV, CTX, (int64_t)C * B, T, cfg.heads, h / cfg.heads, true, stream);
        else rc = attention_cls(QKV.base, CTX.base, odt, (int64_t)C * B, T, cfg.heads, h / cfg.heads, stream);
        if (rc) return rc;
      }
      auto cls_rows = [&](svit_epilogue& e) { e.rows_in = 1, e.rows_out = T, e.row_shift = 0; };  // row b -> X row b * T
      {
        svit_epilogue e{};
        e.bias = vec(SVIT_SEG_BO, l);
        e.bias_gs = vec_stride;
        e.residual = X;
        e.residual_gs = xgs;
        cls_rows(e);
        EpiArgs ea = make_epi(&e, X, xgs, SVIT_F32, B, h);
        if ((rc = gemm_dispatch(plan, CTX, 0, cgs, WM, mat(SVIT_SEG_WO, l), mat_stride, C, B, h, h, ea, stream))) return rc;
      }
      {
        Timed t(plan, stream, SVIT_CLS_LAYERNORM, (double)C * B * h * (4.0 + (split ? 4 : es)));
        if ((rc = layernorm(X, xgs, (int64_t)T * h, vec(SVIT_SEG_LN2_G, l), vec(SVIT_SEG_LN2_B, l), vec_stride, Xn, cgs, h, odt,
                            C, B, h, cfg.ln_eps, stream)))
          return rc;
      }
      {
        svit_epilogue e{};
        e.bias = vec(SVIT_SEG_B1, l);
        e.bias_gs = vec_stride;
        e.gelu = 1;
        EpiArgs ea = make_epi(&e, H.base, (int64_t)B * ff, odt, B, ff);
        out_op(ea, H);
        if ((rc = gemm_dispatch(plan, Xn, 0, cgs, WM, mat(SVIT_SEG_W1, l), mat_stride, C, B, ff, h, ea, stream))) return rc;
      }
      {
        svit_epilogue e{};
        e.bias = vec(SVIT_SEG_B2, l);
        e.bias_gs = vec_stride;
        e.residual = X;
        e.residual_gs = xgs;
        cls_rows(e);
        EpiArgs ea = make_epi(&e, X, xgs, SVIT_F32, B, h);
        if ((rc = gemm_dispatch(plan, H, 0, (int64_t)B * ff, WM, mat(SVIT_SEG_W2, l), mat_stride, C, B, h, ff, ea, stream))) return rc;
      }
      continue;
    }
    {
      Timed t(plan, stream, SVIT_CLS_ATTENTION, 4.0 * C * B * (double)T * T * h);
      if (split) rc = attention_split(QKV, CTX, (int64_t)C * B, T, cfg.heads, h / cfg.heads, false, stream);
      else rc = attention(QKV.base, CTX.base, odt, (int64_t)C * B, T, cfg.heads, h / cfg.heads, stream);
      if (rc) return rc;
    }
    {
      svit_epilogue e{};
      e.bias = vec(SVIT_SEG_BO, l);
      e.bias_gs = vec_stride;
      e.residual = X;
      e.residual_gs = xgs;
      EpiArgs ea = make_epi(&e, X, xgs, SVIT_F32, M, h);
      if ((rc = gemm_dispatch(plan, CTX, 0, xgs, WM, mat(SVIT_SEG_WO, l), mat_stride, C, M, h, h, ea, stream))) return rc;
    }
    {
      Timed t(plan, stream, SVIT_CLS_LAYERNORM, (double)C * M * h * (4.0 + (split ? 4 : es)));
      if ((rc = layernorm(X, xgs, h, vec(SVIT_SEG_LN2_G, l), vec(SVIT_SEG_LN2_B, l), vec_stride, Xn, xgs, h, odt, C, M, h,
                          cfg.ln_eps, stream)))
        return rc;
    }
    {
      svit_epilogue e{};
      e.bias = vec(SVIT_SEG_B1, l);
      e.bias_gs = vec_stride;
      e.gelu = 1;
      EpiArgs ea = make_epi(&e, H.base, (int64_t)M * ff, odt, M, ff);
      out_op(ea, H);
      if ((rc = gemm_dispatch(plan, Xn, 0, xgs, WM, mat(SVIT_SEG_W1, l), mat_stride, C, M, ff, h, ea, stream))) return rc;
    }
    {
      svit_epilogue e{};
      e.bias = vec(SVIT_SEG_B2, l);
      e.bias_gs = vec_stride;
      e.residual = X;
      e.residual_gs = xgs;
      EpiArgs ea = make_epi(&e, X, xgs, SVIT_F32, M, h);
      if ((rc = gemm_dispatch(plan, H, 0, (int64_t)M * ff, WM, mat(SVIT_SEG_W2, l), mat_stride, C, M, h, ff, ea, stream))) return rc;
    }
  }
  // ---- head ----
  return head(X, xgs, wvec, vec_stride, L.find(SVIT_SEG_LNF_G), L.find(SVIT_SEG_LNF_B), L.find(SVIT_SEG_HEAD_W),
              L.find(SVIT_SEG_HEAD_B), logits, logits_stride, C, B, T, h, cfg.n_cls, cfg.ln_eps, stream);
}
}  // namespace
}  // namespace svit

extern "C" int svit_gemm(int precision, const void* A, int64_t a_gs, const void* B, int64_t b_gs, void* out,
                         int64_t out_gs, int out_dtype, int G, int M, int N, int K, const svit_epilogue* epi,
                         svit_stream_t stream) {
  using namespace svit;
  SVIT_CHECK_ARG(A && B && out, "svit_gemm: null pointer");
  SVIT_CHECK_ARG(G >= 1 && M >= 1 && N >= 1 && K >= 1, "svit_gemm: bad sizes");
  const int odt = operand_dtype_of(precision);
  SVIT_CHECK_ARG(odt >= 0, "svit_gemm: unknown precision %d", precision);
  SVIT_CHECK_ARG(out_dtype == SVIT_F32 || out_dtype == odt, "svit_gemm: out_dtype must be f32 or the operand dtype");
  const int fmt = format_of_precision(precision);
  EpiArgs ea = make_epi(epi, out, out_gs, out_dtype, M, N);
  const int64_t m_out = epi && epi->rows_in > 0 ? (int64_t)(M / epi->rows_in) * epi->rows_out : M;
  if (fmt != SVIT_FMT_PLAIN && out_dtype != SVIT_F32) set_out_format(ea, fmt, (int64_t)G * m_out * N);
  const Operand a{const_cast<void*>(A), fmt, (int64_t)(a_gs ? G : 1) * M * K}, b{const_cast<void*>(B), fmt, (int64_t)G * N * K};
  const char* env = getenv("SVIT_FORCE_SIMT_GEMM");
  if (precision == SVIT_PREC_F32 || (fmt == SVIT_FMT_PLAIN && env && env[0] == '1'))
    return gemm_simt(odt, A, a_gs, B, b_gs, G, M, N, K, ea, static_cast<cudaStream_t>(stream));
  return gemm_tc(precision, a, 0, a_gs, b, 0, b_gs, G, M, N, K, ea, static_cast<cudaStream_t>(stream));
}

extern "C" int svit_gemm_ext(int precision, const void* A, int64_t a_gs, const void* B, int64_t b_gs, const void* Ae,
                             const void* Be, void* out, int64_t out_gs, int out_dtype, int G, int M, int N, int K,
                             const svit_epilogue* epi, svit_stream_t stream) {
  using namespace svit;
  SVIT_CHECK_ARG(A && B && Ae && Be && out, "svit_gemm_ext: null pointer");
  SVIT_CHECK_ARG(G >= 1 && M >= 1 && N >= 1 && K >= 1, "svit_gemm_ext: bad sizes");
  const int odt = operand_dtype_of(precision);
  SVIT_CHECK_ARG(odt >= 0 && precision != SVIT_PREC_F32, "svit_gemm_ext: a tensor-core precision is required");
  SVIT_CHECK_ARG(out_dtype == SVIT_F32 || out_dtype == odt, "svit_gemm_ext: out_dtype must be f32 or the operand dtype");
  const int fmt = format_of_precision(precision);
  EpiArgs ea = make_epi(epi, out, out_gs, out_dtype, M, N);
  if (fmt != SVIT_FMT_PLAIN && out_dtype != SVIT_F32) set_out_format(ea, fmt, (int64_t)G * M * N);
  const Operand a{const_cast<void*>(A), fmt, (int64_t)(a_gs ? G : 1) * M * K}, b{const_cast<void*>(B), fmt, (int64_t)(b_gs ? G : 1) * N * K};
  const GemmExt ext{Operand{const_cast<void*>(Ae), fmt, (int64_t)G * M * kGemmExtK}, 0, (int64_t)M * kGemmExtK,
                    Operand{const_cast<void*>(Be), fmt, (int64_t)G * N * kGemmExtK}, 0, (int64_t)N * kGemmExtK};
  return gemm_tc(precision, a, 0, a_gs, b, 0, b_gs, G, M, N, K, ea, static_cast<cudaStream_t>(stream), &ext);
}

extern "C" int svit_plan_timing_begin(svit_plan* plan) {
  using namespace svit;
  SVIT_CHECK_ARG(plan != nullptr, "svit_plan_timing_begin: plan is null");
  plan->timing = true;
  plan->spans.clear();
  plan->pool_used = 0;
  return SVIT_OK;
}

extern "C" int svit_plan_timing_end(svit_plan* plan, svit_timing* out) {
  using namespace svit;
  SVIT_CHECK_ARG(plan && out, "svit_plan_timing_end: null pointer");
  plan->timing = false;
  for (int c = 0; c < SVIT_CLS_COUNT; ++c) out->ms[c] = 0.0, out->work[c] = 0.0, out->launches[c] = 0;
  for (const auto& sp : plan->spans) {
    SVIT_CUDA(cudaEventSynchronize(sp.b));
    float ms = 0.f;
    SVIT_CUDA(cudaEventElapsedTime(&ms, sp.a, sp.b));
    out->ms[sp.cls] += ms;
    out->work[sp.cls] += sp.work;
    out->launches[sp.cls] += 1;
  }
  plan->spans.clear();
  plan->pool_used = 0;
  return SVIT_OK;
}
