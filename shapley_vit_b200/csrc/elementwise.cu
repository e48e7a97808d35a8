// Bandwidth-bound pieces of the batched ViT forward: patch extraction, LayerNorm,
// the [CLS] row of the embedding, and the classifier head (final LayerNorm on the CLS token +
// Linear).  They restate, per coalition group g, what HF's modeling_vit.py does in
// embeddings :100-128, layernorm_before/after :325-326, final layernorm :416 and the
// classifier :641-642 for the model the reference evaluates (federated_learning/utils.py:886).
#include <cstdlib>

#include "elementwise.h"

namespace svit {
namespace {

// Where a kernel's 16-bit / fp32 results go: a plain array of T, or the planes of a split-format packed
// operand array (include/svit.h).  i = element index, a multiple of 4.
template <typename T>
struct PlainOut {
  T* y;
  __device__ __forceinline__ void st4(int64_t i, float a, float b, float c, float d) const { store4<T>(y + i, a, b, c, d); }
};
template <int FMT>
struct SplitOut {
  Operand o;
  __device__ __forceinline__ void st4(int64_t i, float a, float b, float c, float d) const { store4_planes<FMT>(o, i, a, b, c, d); }
};

// ---- patchify: NCHW fp32 images -> [n * np, C*ps*ps] rows in the operand format ------------
template <class Out>
__global__ void patchify_kernel(const float* __restrict__ img, const Out out, int64_t off, int64_t n, int C, int H, int ps) {
  const int gw = H / ps;                       // patches per side
  const int pd = C * ps * ps;                  // row length
  const int q_per_row = pd / 4;                // float4 groups per output row
  const int64_t total = n * gw * gw * (int64_t)q_per_row;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(i % q_per_row);
    const int64_t row = i / q_per_row;
    const int col = q * 4;
    const int c = col / (ps * ps), ky = (col / ps) % ps, kx = col % ps;
    const int64_t im = row / (gw * gw);
    const int p = (int)(row % (gw * gw)), py = p / gw, px = p % gw;
    const float4 v = *reinterpret_cast<const float4*>(
        img + ((im * C + c) * H + (py * ps + ky)) * (int64_t)H + px * ps + kx);
    out.st4(off + row * pd + col, v.x, v.y, v.z, v.w);
  }
}

// ---- LayerNorm: one warp per row, fp32 statistics, two-pass variance ----------------------
constexpr int kLnMaxVec = 8;  // h <= 1024

template <class Out>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, int64_t x_gs, int64_t x_ld,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, int64_t param_gs,
                                                        const Out y, int64_t y_gs, int64_t y_ld, int64_t rows,
                                                        int64_t total_rows, int h, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= total_rows) return;
  const int64_t g = r / rows, rr = r % rows;
  const float* xr = x + g * x_gs + rr * x_ld;
  const int h4 = h >> 2;
  float4 v[kLnMaxVec];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int idx = lane + i * 32;
    if (idx < h4) {
      v[i] = reinterpret_cast<const float4*>(xr)[idx];
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mean = warp_sum(sum) / (float)h;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    if (lane + i * 32 < h4) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      sq += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(sq) / (float)h + eps);
  const float* gm = gamma + g * param_gs;
  const float* bt = beta + g * param_gs;
  const int64_t yr = g * y_gs + rr * y_ld;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int idx = lane + i * 32;
    if (idx < h4) {
      const float4 gg = reinterpret_cast<const float4*>(gm)[idx], bb = reinterpret_cast<const float4*>(bt)[idx];
      y.st4(yr + idx * 4, (v[i].x - mean) * rstd * gg.x + bb.x, (v[i].y - mean) * rstd * gg.y + bb.y,
            (v[i].z - mean) * rstd * gg.z + bb.z, (v[i].w - mean) * rstd * gg.w + bb.w);
    }
  }
}

// Same arithmetic, in the same order, for the hidden sizes that fill whole warps (h = 128 * VPL: 768, 1024):
// persistent warps (one resident wave), no bounds predicates, the (coalition, row) pair advanced
// incrementally instead of by a 64-bit division per row.  Fewer instructions per row than the generic
// kernel -- which matters inside a forward step, where the GEMMs hold the SM clock at the power cap and
// this kernel's issue rate, not HBM, paces it.
template <class Out, int VPL>
__global__ void __launch_bounds__(256) layernorm_fixed_kernel(const float* __restrict__ x, int64_t x_gs, int64_t x_ld,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, int64_t param_gs,
                                                              const Out y, int64_t y_gs, int64_t y_ld, int64_t rows,
                                                              int64_t total_rows, float eps) {
  constexpr int H = VPL * 128;
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  int64_t g = r / rows, rr = r - g * rows;
  for (; r < total_rows; r += nwarps) {
    const float4* xr = reinterpret_cast<const float4*>(x + g * x_gs + rr * x_ld) + lane;
    float4 v[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) v[i] = xr[i * 32];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(sum) / (float)H;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      sq += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(sq) / (float)H + eps);
    const float4* gm = reinterpret_cast<const float4*>(gamma + g * param_gs) + lane;  // L1 hits: shared by the coalition's rows
    const float4* bt = reinterpret_cast<const float4*>(beta + g * param_gs) + lane;
    const int64_t yr = g * y_gs + rr * y_ld + lane * 4;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const float4 gg = gm[i * 32], bb = bt[i * 32];
      y.st4(yr + i * 128, (v[i].x - mean) * rstd * gg.x + bb.x, (v[i].y - mean) * rstd * gg.y + bb.y,
            (v[i].z - mean) * rstd * gg.z + bb.z, (v[i].w - mean) * rstd * gg.w + bb.w);
    }
    rr += nwarps;
    while (rr >= rows) rr -= rows, ++g;
  }
}

// ---- [CLS] rows: X[g, b*T, :] = cls[g] + pos[g][0, :] -------------------------------------
__global__ void embed_cls_kernel(float* __restrict__ X, int64_t x_gs, const float* __restrict__ wvec,
                                 int64_t vec_stride, int64_t off_cls, int64_t off_pos, int B, int T, int h) {
  const int g = blockIdx.y;
  const float* cls = wvec + (size_t)g * vec_stride + off_cls;
  const float* pos = wvec + (size_t)g * vec_stride + off_pos;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)B * h;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / h), k = (int)(i % h);
    X[(size_t)g * x_gs + (size_t)b * T * h + k] = cls[k] + pos[k];
  }
}

// ---- head: final LayerNorm on the CLS token + classifier, one warp per (g, image) ----------
__global__ void __launch_bounds__(256) head_kernel(const float* __restrict__ X, int64_t x_gs,
                                                   const float* __restrict__ wvec, int64_t vec_stride, int64_t off_g,
                                                   int64_t off_b, int64_t off_w, int64_t off_hb,
                                                   float* __restrict__ logits, int64_t logits_stride, int G, int B, int T,
                                                   int h, int n_cls, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= (int64_t)G * B) return;
  const int g = (int)(r / B), b = (int)(r % B);
  const float* xr = X + (size_t)g * x_gs + (size_t)b * T * h;
  const float* wv = wvec + (size_t)g * vec_stride;
  const int h4 = h >> 2;
  float4 v[kLnMaxVec];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int idx = lane + i * 32;
    if (idx < h4) {
      v[i] = reinterpret_cast<const float4*>(xr)[idx];
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mean = warp_sum(sum) / (float)h;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    if (lane + i * 32 < h4) {
      const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      sq += (a * a + bb * bb) + (c * c + d * d);
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(sq) / (float)h + eps);
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int idx = lane + i * 32;
    if (idx < h4) {
      const float4 gg = reinterpret_cast<const float4*>(wv + off_g)[idx];
      const float4 be = reinterpret_cast<const float4*>(wv + off_b)[idx];
      v[i].x = (v[i].x - mean) * rstd * gg.x + be.x;
      v[i].y = (v[i].y - mean) * rstd * gg.y + be.y;
      v[i].z = (v[i].z - mean) * rstd * gg.z + be.z;
      v[i].w = (v[i].w - mean) * rstd * gg.w + be.w;
    }
  }
  for (int k = 0; k < n_cls; ++k) {
    const float* wk = wv + off_w + (size_t)k * h;
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxVec; ++i) {
      const int idx = lane + i * 32;
      if (idx < h4) {
        const float4 ww = reinterpret_cast<const float4*>(wk)[idx];
        dot += (v[i].x * ww.x + v[i].y * ww.y) + (v[i].z * ww.z + v[i].w * ww.w);
      }
    }
    dot = warp_sum(dot);
    if (lane == 0) logits[(size_t)g * logits_stride + (size_t)b * n_cls + k] = dot + wv[off_hb + k];
  }
}

}  // namespace

// ---- fp32 -> split-format packed operand array (SVIT_FMT_X3 / SVIT_FMT_C8) -----------------------
template <int FMT>
__global__ void __launch_bounds__(256) split_operand_kernel(const float* __restrict__ in, const Operand out, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(in) + i);
    store4_planes<FMT>(out, 4 * i, a.x, a.y, a.z, a.w);
  }
}

// run `body(out_policy)` with the output policy of (dtype, operand format)
#define SVIT_WITH_OUT(y, body)                                                              \
  do {                                                                                      \
    if ((y).fmt == SVIT_FMT_X3) {                                                           \
      const SplitOut<SVIT_FMT_X3> out_{(y)};                                                \
      body;                                                                                 \
    } else if ((y).fmt == SVIT_FMT_C8) {                                                    \
      const SplitOut<SVIT_FMT_C8> out_{(y)};                                                \
      body;                                                                                 \
    } else if (out_dtype == SVIT_F32) {                                                     \
      const PlainOut<float> out_{static_cast<float*>((y).base)};                            \
      body;                                                                                 \
    } else if (out_dtype == SVIT_BF16) {                                                    \
      const PlainOut<__nv_bfloat16> out_{static_cast<__nv_bfloat16*>((y).base)};            \
      body;                                                                                 \
    } else if (out_dtype == SVIT_F16) {                                                     \
      const PlainOut<__half> out_{static_cast<__half*>((y).base)};                          \
      body;                                                                                 \
    } else {                                                                                \
      SVIT_FAIL(SVIT_ERR_ARG, "bad output dtype %d", out_dtype);                            \
    }                                                                                       \
  } while (0)

template <class Out>
void launch_patchify(const Out& out, int grid, int block, const float* images, int64_t off, int64_t n, int C, int H, int ps,
                     cudaStream_t stream) {
  patchify_kernel<Out><<<grid, block, 0, stream>>>(images, out, off, n, C, H, ps);
}

// rows start at element offset `off` of the patch matrix
int patchify_at(int out_dtype, const float* images, const Operand& patches, int64_t off, int64_t n, int C, int H, int ps,
                cudaStream_t stream) {
  if (n == 0) return SVIT_OK;
  SVIT_CHECK_ARG(ps % 4 == 0 && H % ps == 0, "patchify: patch must be a multiple of 4 and divide the image");
  const int64_t total = n * (H / ps) * (H / ps) * (int64_t)(C * ps * ps / 4);
  const int block = 256;
  const int grid = (int)std::min<int64_t>((total + block - 1) / block, (int64_t)sm_count() * 16);
  SVIT_WITH_OUT(patches, launch_patchify(out_, grid, block, images, off, n, C, H, ps, stream));
  SVIT_LAUNCH_CHECK("patchify_kernel");
  return SVIT_OK;
}

int split_operand(const float* in, const Operand& out, int64_t elems, cudaStream_t stream) {
  SVIT_CHECK_ARG(elems % 4 == 0 && out.alloc >= elems && out.alloc % 16 == 0, "split_operand: elems %% 4 and alloc %% 16 must be 0, alloc >= elems");
  SVIT_CHECK_ARG(out.fmt == SVIT_FMT_X3 || out.fmt == SVIT_FMT_C8, "split_operand: out must be a split format");
  if (elems == 0) return SVIT_OK;
  const int64_t n4 = elems / 4;
  const int block = 256;
  const int grid = (int)std::min<int64_t>((n4 + block - 1) / block, (int64_t)sm_count() * 16);
  if (out.fmt == SVIT_FMT_X3) split_operand_kernel<SVIT_FMT_X3><<<grid, block, 0, stream>>>(in, out, n4);
  else split_operand_kernel<SVIT_FMT_C8><<<grid, block, 0, stream>>>(in, out, n4);
  SVIT_LAUNCH_CHECK("split_operand_kernel");
  return SVIT_OK;
}

template <class Out>
int launch_layernorm(const Out& out, const float* x, int64_t x_gs, int64_t x_ld, const float* gamma, const float* beta,
                     int64_t param_gs, int64_t y_gs, int64_t y_ld, int G, int64_t rows, int h, float eps, cudaStream_t stream) {
  const int64_t total = (int64_t)G * rows;
  const int wpb = 8;
  static const bool generic_only = [] {  // SVIT_LN_GENERIC=1: the bounds-checked kernel for every h (A/B)
    const char* e = getenv("SVIT_LN_GENERIC");
    return e && e[0] == '1';
  }();
  if ((h == 768 || h == 1024) && !generic_only) {  // whole warps: the persistent fixed-size kernel
    int per_sm = 0;  // one resident wave of persistent warps
    if (h == 768) {
      SVIT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, layernorm_fixed_kernel<Out, 6>, wpb * 32, 0));
    } else {
      SVIT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, layernorm_fixed_kernel<Out, 8>, wpb * 32, 0));
    }
    if (per_sm < 1) per_sm = 1;
    const unsigned pgrid = (unsigned)std::min<int64_t>((total + wpb - 1) / wpb, (int64_t)sm_count() * per_sm);
    if (h == 768)
      layernorm_fixed_kernel<Out, 6><<<pgrid, wpb * 32, 0, stream>>>(x, x_gs, x_ld, gamma, beta, param_gs, out, y_gs, y_ld, rows, total, eps);
    else
      layernorm_fixed_kernel<Out, 8><<<pgrid, wpb * 32, 0, stream>>>(x, x_gs, x_ld, gamma, beta, param_gs, out, y_gs, y_ld, rows, total, eps);
    SVIT_LAUNCH_CHECK("layernorm_fixed_kernel");
    return SVIT_OK;
  }
  const unsigned grid = (unsigned)((total + wpb - 1) / wpb);
  layernorm_kernel<Out><<<grid, wpb * 32, 0, stream>>>(x, x_gs, x_ld, gamma, beta, param_gs, out, y_gs, y_ld, rows, total, h, eps);
  SVIT_LAUNCH_CHECK("layernorm_kernel");
  return SVIT_OK;
}

int layernorm(const float* x, int64_t x_gs, int64_t x_ld, const float* gamma, const float* beta, int64_t param_gs,
              const Operand& y, int64_t y_gs, int64_t y_ld, int out_dtype, int G, int64_t rows, int h, float eps,
              cudaStream_t stream) {
  SVIT_CHECK_ARG(h % 4 == 0 && h <= kLnMaxVec * 128, "layernorm: h=%d must be a multiple of 4 and <= 1024", h);
  SVIT_CHECK_ARG(x_ld % 4 == 0 && y_ld % 4 == 0 && x_gs % 4 == 0 && y_gs % 4 == 0 && param_gs % 4 == 0, "layernorm: strides must be multiples of 4");
  if ((int64_t)G * rows == 0) return SVIT_OK;
  int rc = SVIT_OK;
  SVIT_WITH_OUT(y, rc = launch_layernorm(out_, x, x_gs, x_ld, gamma, beta, param_gs, y_gs, y_ld, G, rows, h, eps, stream));
  return rc;
}

int embed_cls(float* X, int64_t x_gs, const float* wvec, int64_t vec_stride, int64_t off_cls, int64_t off_pos, int G,
              int B, int T, int h, cudaStream_t stream) {
  dim3 grid((unsigned)std::min<int64_t>(((int64_t)B * h + 255) / 256, 1024), G);
  embed_cls_kernel<<<grid, 256, 0, stream>>>(X, x_gs, wvec, vec_stride, off_cls, off_pos, B, T, h);
  SVIT_LAUNCH_CHECK("embed_cls_kernel");
  return SVIT_OK;
}

int head(const float* X, int64_t x_gs, const float* wvec, int64_t vec_stride, int64_t off_g, int64_t off_b,
         int64_t off_w, int64_t off_hb, float* logits, int64_t logits_stride, int G, int B, int T, int h, int n_cls,
         float eps, cudaStream_t stream) {
  SVIT_CHECK_ARG(h % 4 == 0 && h <= kLnMaxVec * 128, "head: h=%d unsupported", h);
  const int64_t total = (int64_t)G * B;
  const int wpb = 8;
  head_kernel<<<(unsigned)((total + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
      X, x_gs, wvec, vec_stride, off_g, off_b, off_w, off_hb, logits, logits_stride, G, B, T, h, n_cls, eps);
  SVIT_LAUNCH_CHECK("head_kernel");
  return SVIT_OK;
}

}  // namespace svit

extern "C" int svit_layernorm(const float* x, int64_t x_gs, int64_t x_ld, const float* gamma, const float* beta,
                              int64_t param_gs, void* y, int64_t y_gs, int64_t y_ld, int out_dtype, int out_fmt,
                              int64_t out_alloc, int G, int64_t rows, int h, float eps, svit_stream_t stream) {
  using namespace svit;
  SVIT_CHECK_ARG(x && gamma && beta && y, "svit_layernorm: null pointer");
  SVIT_CHECK_ARG(G >= 1 && rows >= 0, "svit_layernorm: bad sizes");
  SVIT_CHECK_ARG(out_fmt == SVIT_FMT_PLAIN || out_alloc % 16 == 0, "svit_layernorm: plane pitch must be a multiple of 16");
  if (!aligned16(x) || !aligned16(gamma) || !aligned16(beta) || !aligned16(y))
    SVIT_FAIL(SVIT_ERR_ALIGN, "svit_layernorm: pointers must be 16-byte aligned");
  return layernorm(x, x_gs, x_ld, gamma, beta, param_gs, Operand{y, out_fmt, out_alloc}, y_gs, y_ld, out_dtype, G, rows, h, eps,
                   static_cast<cudaStream_t>(stream));
}

extern "C" int svit_split_operand(const float* in, void* out, int64_t out_alloc, int out_fmt, int64_t elems,
                                  svit_stream_t stream) {
  using namespace svit;
  SVIT_CHECK_ARG(in && out && elems >= 0, "svit_split_operand: bad arguments");
  if (!aligned16(in) || !aligned16(out)) SVIT_FAIL(SVIT_ERR_ALIGN, "svit_split_operand: pointers must be 16-byte aligned");
  return split_operand(in, Operand{out, out_fmt, out_alloc}, elems, static_cast<cudaStream_t>(stream));
}
