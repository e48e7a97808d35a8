// Fused-softmax attention for short, non-causal sequences (ViT: T = 197 at 224 px, 5 at 32 px).
//
// Restates HF ViTSelfAttention (modeling_vit.py:171-196, 220-252) as evaluated by the
// reference (federated_learning/utils.py:886): per (sequence, head)
//     ctx = softmax(q k^T * d^-0.5) v,   no mask, no dropout in eval.
// The whole K and V of one head live in shared memory (T*d*4*2 bytes = 101 KB for T=197, d=64),
// a warp owns one query row at a time: scores (7 keys per lane) -> fp32 softmax in registers ->
// P V with two output channels per lane.  Scores never touch HBM.
//
// This is the fp32 CUDA-core kernel: the exact path for fp32-stored activations (SVIT_PREC_F32 /
// TF32) and for head sizes other than 64; 16-bit operands with head_dim 64 take the tensor-core
// kernel in attention_mma.cu.  Operands may be stored in fp32 / bf16 / fp16, all arithmetic is fp32.
#include <cstdlib>

#include "elementwise.h"

namespace svit {
namespace {

constexpr int kAttWarps = 8;

// dynamic smem: Ks [T][D+1] | Vs [T][D] | per-warp q [D] and p [Tpad]
template <typename T, int D>
__global__ void __launch_bounds__(kAttWarps * 32) attention_kernel(const T* __restrict__ qkv, T* __restrict__ ctx,
                                                                   int Tn, int heads) {
  extern __shared__ float att_smem[];
  const int h = heads * D;
  const int64_t seq = blockIdx.x;
  const int head = blockIdx.y;
  const int Tpad = (Tn + 31) & ~31;
  float* Ks = att_smem;
  float* Vs = Ks + (size_t)Tn * (D + 1);
  float* wq = Vs + (size_t)Tn * D;
  float* wp = wq + kAttWarps * D;
  const T* base = qkv + seq * (int64_t)Tn * 3 * h;
  for (int i = threadIdx.x; i < Tn * D; i += blockDim.x) {
    const int t = i / D, d = i % D;
    Ks[t * (D + 1) + d] = Cvt<T>::to_f(base[(size_t)t * 3 * h + h + head * D + d]);
    Vs[t * D + d] = Cvt<T>::to_f(base[(size_t)t * 3 * h + 2 * h + head * D + d]);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* q = wq + warp * D;
  float* p = wp + warp * Tpad;
  const float scale = rsqrtf((float)D);
  constexpr int KPL = 8;  // keys per lane: supports T <= 256
  for (int t = warp; t < Tn; t += kAttWarps) {
    for (int d = lane; d < D; d += 32) q[d] = Cvt<T>::to_f(base[(size_t)t * 3 * h + head * D + d]);
    __syncwarp();
    float s[KPL];
#pragma unroll
    for (int k = 0; k < KPL; ++k) s[k] = 0.f;
    for (int d = 0; d < D; ++d) {
      const float qd = q[d];
#pragma unroll
      for (int k = 0; k < KPL; ++k) {
        const int j = lane + k * 32;
        if (j < Tn) s[k] = fmaf(qd, Ks[j * (D + 1) + d], s[k]);
      }
    }
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < KPL; ++k) {
      s[k] *= scale;
      if (lane + k * 32 < Tn) m = fmaxf(m, s[k]);
    }
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < KPL; ++k) {
      const int j = lane + k * 32;
      s[k] = j < Tn ? expf(s[k] - m) : 0.f;
      sum += s[k];
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int k = 0; k < KPL; ++k) {
      const int j = lane + k * 32;
      if (j < Tpad) p[j] = s[k] * inv;
    }
    __syncwarp();
    // P V: lane owns channels lane, lane+32, ...
    float o[D / 32];
#pragma unroll
    for (int c = 0; c < D / 32; ++c) o[c] = 0.f;
    for (int j = 0; j < Tn; ++j) {
      const float pj = p[j];
#pragma unroll
      for (int c = 0; c < D / 32; ++c) o[c] = fmaf(pj, Vs[j * D + lane + c * 32], o[c]);
    }
    T* out = ctx + (seq * (int64_t)Tn + t) * h + head * D;
#pragma unroll
    for (int c = 0; c < D / 32; ++c) out[lane + c * 32] = Cvt<T>::from_f(o[c]);
    __syncwarp();
  }
}

template <typename T, int D>
int launch_att(const void* qkv, void* ctx, int64_t n_seq, int Tn, int heads, cudaStream_t stream) {
  const int Tpad = (Tn + 31) & ~31;
  const size_t smem = ((size_t)Tn * (D + 1) + (size_t)Tn * D + kAttWarps * D + kAttWarps * Tpad) * sizeof(float);
  SVIT_CHECK_ARG(smem <= 227 * 1024, "attention: T=%d does not fit shared memory", Tn);
  auto kern = attention_kernel<T, D>;
  SVIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SVIT_CHECK_ARG(n_seq <= 0x7fffffff, "attention: too many sequences");
  dim3 grid((unsigned)n_seq, heads);
  kern<<<grid, kAttWarps * 32, smem, stream>>>((const T*)qkv, (T*)ctx, Tn, heads);
  SVIT_LAUNCH_CHECK("attention_kernel");
  return SVIT_OK;
}

// ---- register-tiled fp32 kernel for head_dim 64 ---------------------------------------------------
// Same arithmetic as attention_kernel (one fp32 accumulator per score, keys and channels summed in
// index order, expf, p = e * (1 / sum)), so the results are bit-identical to it -- but a warp owns
// EIGHT query rows at a time: every K element read from shared memory feeds 8 FMAs and every V
// element 4, which lifts the kernel from the shared-memory load rate (one load per FMA, 10 TFLOP/s)
// to the FMA rate.  This is the attention of the fp32-storage precisions (f32, tf32, f16x3).
// Shared memory: Ks [TP][68] | Vs [TP][64] | per warp q [8][64] and p [8][TP + 4]; TP = 32 * KPL.
constexpr int kRtRows = 8;

template <typename T>
__device__ __forceinline__ float4 load4(const T* p) {
  return make_float4(Cvt<T>::to_f(p[0]), Cvt<T>::to_f(p[1]), Cvt<T>::to_f(p[2]), Cvt<T>::to_f(p[3]));
}
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
template <typename T>
__device__ __forceinline__ void store4(T* p, float4 v) {
  p[0] = Cvt<T>::from_f(v.x), p[1] = Cvt<T>::from_f(v.y), p[2] = Cvt<T>::from_f(v.z), p[3] = Cvt<T>::from_f(v.w);
}
template <>
__device__ __forceinline__ void store4<float>(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}

template <typename T, int KPL>
__global__ void __launch_bounds__(kAttWarps * 32) attention_rt_kernel(const T* __restrict__ qkv, T* __restrict__ ctx,
                                                                      int Tn, int heads) {
  constexpr int D = 64, KS = D + 4, TP = KPL * 32, PS = TP + 4;
  extern __shared__ __align__(16) float att_smem[];
  float* Ks = att_smem;
  float* Vs = Ks + TP * KS;
  float* Qs = Vs + TP * D;
  float* Ps = Qs + kAttWarps * kRtRows * D;
  const int h = heads * D;
  const int64_t seq = blockIdx.x;
  const int head = blockIdx.y;
  const T* base = qkv + seq * (int64_t)Tn * 3 * h + head * D;
  for (int i = threadIdx.x; i < TP * (D / 4); i += blockDim.x) {  // rows >= T are zero: no masks in the loops below
    const int t = i / (D / 4), c = (i % (D / 4)) * 4;
    float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
    if (t < Tn) {
      kv = load4<T>(base + (size_t)t * 3 * h + h + c);
      vv = load4<T>(base + (size_t)t * 3 * h + 2 * h + c);
    }
    *reinterpret_cast<float4*>(Ks + t * KS + c) = kv;
    *reinterpret_cast<float4*>(Vs + t * D + c) = vv;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* qw = Qs + warp * kRtRows * D;
  float* pw = Ps + warp * kRtRows * PS;
  const float scale = rsqrtf((float)D);
  for (int t0 = warp * kRtRows; t0 < Tn; t0 += kAttWarps * kRtRows) {
#pragma unroll
    for (int r = 0; r < kRtRows * D / 4 / 32; ++r) {  // the 8 query rows (zero past the end)
      const int idx = lane + 32 * r, row = idx / (D / 4), c = (idx % (D / 4)) * 4;
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t0 + row < Tn) q = load4<T>(base + (size_t)(t0 + row) * 3 * h + c);
      *reinterpret_cast<float4*>(qw + row * D + c) = q;
    }
    __syncwarp();
    // ---- scores: lane owns keys lane + 32 k ----
    float s[kRtRows][KPL];
#pragma unroll
    for (int i = 0; i < kRtRows; ++i)
#pragma unroll
      for (int k = 0; k < KPL; ++k) s[i][k] = 0.f;
#pragma unroll 2
    for (int d = 0; d < D; d += 4) {
      float4 q[kRtRows];
#pragma unroll
      for (int i = 0; i < kRtRows; ++i) q[i] = *reinterpret_cast<const float4*>(qw + i * D + d);
#pragma unroll
      for (int k = 0; k < KPL; ++k) {
        const float4 kk = *reinterpret_cast<const float4*>(Ks + (lane + 32 * k) * KS + d);
#pragma unroll
        for (int i = 0; i < kRtRows; ++i) {
          s[i][k] = fmaf(q[i].x, kk.x, s[i][k]);
          s[i][k] = fmaf(q[i].y, kk.y, s[i][k]);
          s[i][k] = fmaf(q[i].z, kk.z, s[i][k]);
          s[i][k] = fmaf(q[i].w, kk.w, s[i][k]);
        }
      }
    }
    // ---- softmax per row, probabilities to shared memory ----
#pragma unroll
    for (int i = 0; i < kRtRows; ++i) {
      float m = -INFINITY;
#pragma unroll
      for (int k = 0; k < KPL; ++k) {
        s[i][k] *= scale;
        if (lane + k * 32 < Tn) m = fmaxf(m, s[i][k]);
      }
      m = warp_max(m);
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < KPL; ++k) {
        s[i][k] = lane + k * 32 < Tn ? expf(s[i][k] - m) : 0.f;
        sum += s[i][k];
      }
      sum = warp_sum(sum);
      const float inv = 1.0f / sum;
#pragma unroll
      for (int k = 0; k < KPL; ++k) pw[i * PS + lane + k * 32] = s[i][k] * inv;
    }
    __syncwarp();
    // ---- P V: half-warp `half` owns rows 4 half .. 4 half + 3, each lane 4 channels ----
    const int half = lane >> 4, c4 = (lane & 15) * 4;
    float4 o[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) o[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* pr = pw + half * 4 * PS;
    const int jend = (Tn + 3) & ~3;
    for (int j4 = 0; j4 < jend; j4 += 4) {
      float4 p[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) p[r] = *reinterpret_cast<const float4*>(pr + r * PS + j4);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float4 v = *reinterpret_cast<const float4*>(Vs + (j4 + jj) * D + c4);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float pj = jj == 0 ? p[r].x : jj == 1 ? p[r].y : jj == 2 ? p[r].z : p[r].w;
          o[r].x = fmaf(pj, v.x, o[r].x);
          o[r].y = fmaf(pj, v.y, o[r].y);
          o[r].z = fmaf(pj, v.z, o[r].z);
          o[r].w = fmaf(pj, v.w, o[r].w);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = t0 + half * 4 + r;
      if (row < Tn) store4<T>(ctx + (seq * (int64_t)Tn + row) * h + head * D + c4, o[r]);
    }
    __syncwarp();
  }
}

template <typename T, int KPL>
int launch_rt(const void* qkv, void* ctx, int64_t n_seq, int Tn, int heads, cudaStream_t stream) {
  constexpr int D = 64, TP = KPL * 32;
  const size_t smem = ((size_t)TP * (D + 4) + (size_t)TP * D + kAttWarps * kRtRows * D + kAttWarps * kRtRows * (TP + 4)) * sizeof(float);
  auto kern = attention_rt_kernel<T, KPL>;
  SVIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SVIT_CHECK_ARG(n_seq <= 0x7fffffff, "attention: too many sequences");
  dim3 grid((unsigned)n_seq, heads);
  kern<<<grid, kAttWarps * 32, smem, stream>>>((const T*)qkv, (T*)ctx, Tn, heads);
  SVIT_LAUNCH_CHECK("attention_rt_kernel");
  return SVIT_OK;
}

template <typename T>
int dispatch_rt(const void* qkv, void* ctx, int64_t n_seq, int Tn, int heads, cudaStream_t stream) {
  switch ((Tn + 31) / 32) {
    case 1: return launch_rt<T, 1>(qkv, ctx, n_seq, Tn, heads, stream);
    case 2: return launch_rt<T, 2>(qkv, ctx, n_seq, Tn, heads, stream);
    case 3: return launch_rt<T, 3>(qkv, ctx, n_seq, Tn, heads, stream);
    case 4: return launch_rt<T, 4>(qkv, ctx, n_seq, Tn, heads, stream);
    case 5: return launch_rt<T, 5>(qkv, ctx, n_seq, Tn, heads, stream);
    case 6: return launch_rt<T, 6>(qkv, ctx, n_seq, Tn, heads, stream);
    case 7: return launch_rt<T, 7>(qkv, ctx, n_seq, Tn, heads, stream);
    default: return launch_rt<T, 8>(qkv, ctx, n_seq, Tn, heads, stream);
  }
}

template <typename T>
int dispatch_d(const void* qkv, void* ctx, int64_t n_seq, int Tn, int heads, int head_dim, cudaStream_t stream) {
  static const bool no_rt = [] {  // SVIT_ATTENTION_NO_RT=1: the one-row-per-warp kernel for every head size (A/B)
    const char* e = getenv("SVIT_ATTENTION_NO_RT");
    return e && e[0] == '1';
  }();
  if (head_dim == 64 && (heads * head_dim) % 4 == 0 && !no_rt) return dispatch_rt<T>(qkv, ctx, n_seq, Tn, heads, stream);
  switch (head_dim) {
    case 32: return launch_att<T, 32>(qkv, ctx, n_seq, Tn, heads, stream);
    case 64: return launch_att<T, 64>(qkv, ctx, n_seq, Tn, heads, stream);
    case 96: return launch_att<T, 96>(qkv, ctx, n_seq, Tn, heads, stream);
    case 128: return launch_att<T, 128>(qkv, ctx, n_seq, Tn, heads, stream);
    default: SVIT_FAIL(SVIT_ERR_UNSUPPORTED, "attention: head_dim %d not supported (32/64/96/128)", head_dim);
  }
}


// ---- [CLS]-query attention: the last encoder layer ---------------------------------------------
// The classifier reads only the [CLS] token (HF modeling_vit.py:641-642), so in the LAST layer only
// the [CLS] query row has to attend (keys and values still come from every token).  One warp per
// (sequence, head): lanes split the keys for the scores and the softmax, then split the channels
// for P V.  ctx_cls is compact: [n_seq, h].  All arithmetic fp32; operands fp32 / bf16 / fp16.
// dot product of a key row (D elements of T, 16-byte aligned) with q (fp32, shared memory), 16-byte loads
template <typename T, int D>
__device__ __forceinline__ float dot_row(const T* __restrict__ row, const float* __restrict__ q) {
  constexpr int EPV = 16 / sizeof(T);  // elements per 16-byte load
  float acc = 0.f;
#pragma unroll
  for (int v = 0; v < D / EPV; ++v) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(row) + v);
    const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
    for (int i = 0; i < EPV; ++i) acc = fmaf(q[v * EPV + i], Cvt<T>::to_f(e[i]), acc);
  }
  return acc;
}

template <typename T, int D>
__global__ void __launch_bounds__(256) attention_cls_kernel(const T* __restrict__ qkv, T* __restrict__ ctx_cls, int Tn,
                                                            int heads, int64_t n_items) {
  __shared__ __align__(16) float q_s[8][D];
  __shared__ float p_s[8][256];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t item = blockIdx.x * 8LL + w;
  if (item >= n_items) return;
  const int64_t seq = item / heads;
  const int head = (int)(item % heads);
  const int h = heads * D;
  const T* base = qkv + seq * (int64_t)Tn * 3 * h + head * D;
  for (int d = lane; d < D; d += 32) q_s[w][d] = Cvt<T>::to_f(base[d]);
  __syncwarp();
  constexpr int KPL = 8;  // keys per lane: T <= 256
  float s[KPL];
  const float scale = rsqrtf((float)D);
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < KPL; ++k) {  // every lane reads whole 128-byte key rows (full sectors)
    const int j = lane + k * 32;
    s[k] = -INFINITY;
    if (j < Tn) {
      s[k] = dot_row<T, D>(base + (size_t)j * 3 * h + h, q_s[w]) * scale;
      m = fmaxf(m, s[k]);
    }
  }
  m = warp_max(m);
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < KPL; ++k) {
    const int j = lane + k * 32;
    s[k] = j < Tn ? expf(s[k] - m) : 0.f;
    sum += s[k];
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
#pragma unroll
  for (int k = 0; k < KPL; ++k) p_s[w][lane + k * 32] = s[k] * inv;
  __syncwarp();
  // P V: the lane owns D / 32 CONSECUTIVE channels, so a warp reads each value row as one contiguous segment
  constexpr int CPL = D / 32;
  float o[CPL];
#pragma unroll
  for (int c = 0; c < CPL; ++c) o[c] = 0.f;
  const T* vbase = base + 2 * h + lane * CPL;
#pragma unroll 4
  for (int j = 0; j < Tn; ++j) {
    const float pj = p_s[w][j];
    T v[CPL];
    if constexpr (CPL * sizeof(T) == 4) {
      *reinterpret_cast<uint32_t*>(v) = __ldg(reinterpret_cast<const uint32_t*>(vbase + (size_t)j * 3 * h));
    } else if constexpr (CPL * sizeof(T) == 8) {
      *reinterpret_cast<uint2*>(v) = __ldg(reinterpret_cast<const uint2*>(vbase + (size_t)j * 3 * h));
    } else {
#pragma unroll
      for (int c = 0; c < CPL; ++c) v[c] = vbase[(size_t)j * 3 * h + c];
    }
#pragma unroll
    for (int c = 0; c < CPL; ++c) o[c] = fmaf(pj, Cvt<T>::to_f(v[c]), o[c]);
  }
  T* out = ctx_cls + seq * (int64_t)h + head * D + lane * CPL;
#pragma unroll
  for (int c = 0; c < CPL; ++c) out[c] = Cvt<T>::from_f(o[c]);
}

template <typename T>
int launch_cls(const void* qkv, void* ctx, int64_t n_seq, int Tn, int heads, int head_dim, cudaStream_t stream) {
  const int64_t items = n_seq * heads;
  const unsigned grid = (unsigned)((items + 7) / 8);
  switch (head_dim) {
    case 32: attention_cls_kernel<T, 32><<<grid, 256, 0, stream>>>((const T*)qkv, (T*)ctx, Tn, heads, items); break;
    case 64: attention_cls_kernel<T, 64><<<grid, 256, 0, stream>>>((const T*)qkv, (T*)ctx, Tn, heads, items); break;
    case 96: attention_cls_kernel<T, 96><<<grid, 256, 0, stream>>>((const T*)qkv, (T*)ctx, Tn, heads, items); break;
    case 128: attention_cls_kernel<T, 128><<<grid, 256, 0, stream>>>((const T*)qkv, (T*)ctx, Tn, heads, items); break;
    default: SVIT_FAIL(SVIT_ERR_UNSUPPORTED, "attention: head_dim %d not supported (32/64/96/128)", head_dim);
  }
  SVIT_LAUNCH_CHECK("attention_cls_kernel");
  return SVIT_OK;
}

}  // namespace

// qkv [n_seq, T, 3h] -> ctx_cls [n_seq, h]: attention output of the [CLS] query (token 0) only
int attention_cls(const void* qkv, void* ctx_cls, int dtype, int64_t n_seq, int Tn, int heads, int head_dim,
                  cudaStream_t stream) {
  if (n_seq == 0) return SVIT_OK;
  SVIT_CHECK_ARG(Tn >= 1 && Tn <= 256, "attention: T=%d out of range (1..256)", Tn);
  switch (dtype) {
    case SVIT_F32: return launch_cls<float>(qkv, ctx_cls, n_seq, Tn, heads, head_dim, stream);
    case SVIT_BF16: return launch_cls<__nv_bfloat16>(qkv, ctx_cls, n_seq, Tn, heads, head_dim, stream);
    case SVIT_F16: return launch_cls<__half>(qkv, ctx_cls, n_seq, Tn, heads, head_dim, stream);
    default: SVIT_FAIL(SVIT_ERR_ARG, "attention: bad dtype %d", dtype);
  }
}

int attention(const void* qkv, void* ctx, int dtype, int64_t n_seq, int Tn, int heads, int head_dim,
              cudaStream_t stream) {
  if (n_seq == 0) return SVIT_OK;
  SVIT_CHECK_ARG(Tn >= 1 && Tn <= 256, "attention: T=%d out of range (1..256)", Tn);
  // 16-bit operands with the ViT head size go to the tensor-core kernel (attention_mma.cu)
  static const bool force_simt = [] {
    const char* e = getenv("SVIT_ATTENTION_SIMT");
    return e && e[0] == '1';
  }();
  if (!force_simt && head_dim == 64 && (dtype == SVIT_F16 || dtype == SVIT_BF16) && (heads * head_dim) % 8 == 0) {
    // 128 < T <= 256 (ViT at 224 px): tcgen05 / TMEM kernel; shorter sequences: warp-level MMA kernel
    static const bool no_tc = [] {
      const char* e = getenv("SVIT_ATTENTION_MMA_SYNC");
      return e && e[0] == '1';
    }();
    if (Tn > 128 && !no_tc) return attention_tc(qkv, ctx, dtype, n_seq, Tn, heads, stream);
    return attention_mma(qkv, ctx, dtype, n_seq, Tn, heads, stream);
  }
  switch (dtype) {
    case SVIT_F32: return dispatch_d<float>(qkv, ctx, n_seq, Tn, heads, head_dim, stream);
    case SVIT_BF16: return dispatch_d<__nv_bfloat16>(qkv, ctx, n_seq, Tn, heads, head_dim, stream);
    case SVIT_F16: return dispatch_d<__half>(qkv, ctx, n_seq, Tn, heads, head_dim, stream);
    default: SVIT_FAIL(SVIT_ERR_ARG, "attention: bad dtype %d", dtype);
  }
}

}  // namespace svit

extern "C" int svit_attention(const void* qkv, void* ctx, int dtype, int64_t n_seq, int T, int heads, int head_dim,
                              svit_stream_t stream) {
  using namespace svit;
  SVIT_CHECK_ARG(qkv && ctx, "svit_attention: null pointer");
  SVIT_CHECK_ARG(n_seq >= 0 && heads >= 1, "svit_attention: bad sizes");
  return attention(qkv, ctx, dtype, n_seq, T, heads, head_dim, static_cast<cudaStream_t>(stream));
}

extern "C" int svit_attention_split(const void* qkv, void* ctx, int ctx_fmt, int64_t n_seq, int T, int heads, int head_dim,
                                    svit_stream_t stream) {
  using namespace svit;
  SVIT_CHECK_ARG(qkv && ctx, "svit_attention_split: null pointer");
  SVIT_CHECK_ARG(n_seq >= 0 && heads >= 1, "svit_attention_split: bad sizes");
  const int64_t h = (int64_t)heads * head_dim;
  return attention_split(Operand{const_cast<void*>(qkv), SVIT_FMT_X3, n_seq * T * 3 * h}, Operand{ctx, ctx_fmt, n_seq * T * h},
                         n_seq, T, heads, head_dim, false, static_cast<cudaStream_t>(stream));
}
