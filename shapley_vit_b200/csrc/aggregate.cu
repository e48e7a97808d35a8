// K1 -- coalition aggregation (HBM-bound).
//
// Replaces, for a whole batch of C coalitions at once, the reference's
//   get_aggregated_model      federated_learning/utils.py:781-792   agg = sum_j r_j * Delta_j
//   ServerBase.model_agg_lazy federated_learning/server2.py:121-127 W_S = W_0 + agg
// which the reference runs once per coalition and per state_dict key (2|S|-1 elementwise
// launches per key, each re-reading its operands).
//
// Data movement: the stacked deltas [N, P] and W_0 [P] are streamed from HBM exactly once.
// A persistent CTA owns tiles of TILE consecutive parameters; one elected thread stages the
// N+1 rows of a tile into shared memory with 1-D TMA bulk copies (cp.async.bulk, completion on
// an mbarrier), STAGES tiles deep, so the loads in flight do not depend on occupancy or
// registers.  Each thread owns 4 consecutive parameters, keeps 16 coalition accumulators
// (x4 lanes) in registers, walks the clients in ascending order reading its float4 from shared
// memory, and writes every coalition's row with 64/128-bit streaming stores.
//
// Arithmetic: every product and every sum is a separately rounded fp32 operation
// (__fmul_rn/__fadd_rn; ptxas would otherwise contract to FMA), in ascending client order,
// which is bit-identical to the reference whenever frozenset(coalition) iterates in ascending
// order.  Algorithmic bytes per launch:
//   4*P*(N+1) + sizeof(out)*P*C.
#include "common.cuh"

namespace svit {
namespace {

constexpr int kCChunk = 16;  // coalition accumulators held in registers per pass

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // bounded: a lost TMA completion traps instead of hanging the GPU
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
    if (spin > (1u << 26)) __trap();
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

template <typename OutT> struct Store4;
template <> struct Store4<float> {
  static __device__ __forceinline__ void st(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }
};
template <> struct Store4<__nv_bfloat16> {
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float4 v) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&lo);
    u.y = *reinterpret_cast<uint32_t*>(&hi);
    __stcs(reinterpret_cast<uint2*>(p), u);
  }
};
template <> struct Store4<__half> {
  static __device__ __forceinline__ void st(__half* p, float4 v) {
    __half2 lo = __halves2half2(Cvt<__half>::from_f(v.x), Cvt<__half>::from_f(v.y));
    __half2 hi = __halves2half2(Cvt<__half>::from_f(v.z), Cvt<__half>::from_f(v.w));
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&lo);
    u.y = *reinterpret_cast<uint32_t*>(&hi);
    __stcs(reinterpret_cast<uint2*>(p), u);
  }
};

struct AggParams {
  const float* deltas;
  int64_t delta_stride;
  const float* w0;  // may be null
  const float* ratios;
  void* out;
  int64_t out_stride;
  int64_t P;
  int N, C, stages;
  int64_t num_tiles;
};

// dynamic smem: [STAGES][(N+1)][TILE] floats | ratios [C*N] floats | mbarriers [STAGES]
template <typename OutT, int BLOCK>
__global__ void __launch_bounds__(BLOCK) aggregate_kernel(const AggParams p) {
  constexpr int TILE = BLOCK * 4;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int N = p.N, C = p.C, S = p.stages;
  const int rows = N + 1;  // row N holds W_0
  float* stage_base = reinterpret_cast<float*>(smem_raw);
  float* s_ratio = stage_base + (size_t)S * rows * TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ratio + (((size_t)C * N + 1) & ~(size_t)1));
  const int tid = threadIdx.x;

  for (int i = tid; i < C * N; i += BLOCK) s_ratio[i] = p.ratios[i];
  if (tid == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int64_t P = p.P;
  auto tile_len = [&](int64_t t) -> int { return (int)min((int64_t)TILE, P - t * TILE); };
  // a tile goes through TMA when its length is a multiple of 4 floats (16 B granules)
  auto issue = [&](int64_t t, int s) {  // thread 0 only
    const int len = tile_len(t);
    if (len & 3) return;  // ragged last tile: staged by all threads in the consumer path
    float* dst = stage_base + (size_t)s * rows * TILE;
    const uint32_t bytes = (uint32_t)len * 4u;
    mbar_expect_tx(&bars[s], bytes * (uint32_t)(p.w0 ? rows : N));
    for (int j = 0; j < N; ++j)
      bulk_g2s(dst + (size_t)j * TILE, p.deltas + (size_t)j * p.delta_stride + t * TILE, bytes, &bars[s]);
    if (p.w0) bulk_g2s(dst + (size_t)N * TILE, p.w0 + t * TILE, bytes, &bars[s]);
  };

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      const int64_t t = blockIdx.x + (int64_t)s * gridDim.x;
      if (t < p.num_tiles) issue(t, s);
    }
  }

  int s = 0;
  uint32_t parity = 0;
  for (int64_t t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
    const int len = tile_len(t);
    float* st = stage_base + (size_t)s * rows * TILE;
    if (len & 3) {  // ragged tail (at most one tile per launch): guarded scalar staging
      for (int j = 0; j < rows; ++j) {
        const float* src = j < N ? p.deltas + (size_t)j * p.delta_stride + t * TILE : (p.w0 ? p.w0 + t * TILE : nullptr);
        for (int e = tid; e < TILE; e += BLOCK) st[(size_t)j * TILE + e] = (src && e < len) ? src[e] : 0.f;
      }
      __syncthreads();
    } else {
      mbar_wait(&bars[s], parity);
    }
    const int e0 = tid * 4;
    if (e0 < len) {
      const float4 w = p.w0 ? *reinterpret_cast<const float4*>(st + (size_t)N * TILE + e0) : make_float4(0, 0, 0, 0);
      OutT* outp = reinterpret_cast<OutT*>(p.out) + t * TILE + e0;
      for (int c0 = 0; c0 < C; c0 += kCChunk) {
        float4 acc[kCChunk];
#pragma unroll
        for (int cc = 0; cc < kCChunk; ++cc) acc[cc] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = 0; j < N; ++j) {
          const float4 d = *reinterpret_cast<const float4*>(st + (size_t)j * TILE + e0);
#pragma unroll
          for (int cc = 0; cc < kCChunk; ++cc) {
            if (c0 + cc < C) {
              const float r = s_ratio[(c0 + cc) * N + j];  // warp-uniform broadcast read
              if (r != 0.f) {
                acc[cc].x = __fadd_rn(acc[cc].x, __fmul_rn(r, d.x));
                acc[cc].y = __fadd_rn(acc[cc].y, __fmul_rn(r, d.y));
                acc[cc].z = __fadd_rn(acc[cc].z, __fmul_rn(r, d.z));
                acc[cc].w = __fadd_rn(acc[cc].w, __fmul_rn(r, d.w));
              }
            }
          }
        }
#pragma unroll
        for (int cc = 0; cc < kCChunk; ++cc) {
          if (c0 + cc < C) {
            float4 v;
            v.x = __fadd_rn(w.x, acc[cc].x);
            v.y = __fadd_rn(w.y, acc[cc].y);
            v.z = __fadd_rn(w.z, acc[cc].z);
            v.w = __fadd_rn(w.w, acc[cc].w);
            OutT* o = outp + (size_t)(c0 + cc) * p.out_stride;
            if (e0 + 4 <= len) {
              Store4<OutT>::st(o, v);
            } else {  // ragged tail: element-wise
              const float vv[4] = {v.x, v.y, v.z, v.w};
              for (int q = 0; q < 4 && e0 + q < len; ++q) o[q] = Cvt<OutT>::from_f(vv[q]);
            }
          }
        }
      }
    }
    __syncthreads();  // everyone is done reading stage s
    if (tid == 0) {
      const int64_t tn = t + (int64_t)S * gridDim.x;
      if (tn < p.num_tiles) issue(tn, s);
    }
    if (++s == S) {
      s = 0;
      parity ^= 1;
    }
  }
}

template <typename OutT, int BLOCK>
int launch(const AggParams& base, cudaStream_t stream) {
  AggParams p = base;
  constexpr int TILE = BLOCK * 4;
  p.num_tiles = (p.P + TILE - 1) / TILE;
  const size_t stage_bytes = (size_t)(p.N + 1) * TILE * 4;
  const size_t fixed = ((((size_t)p.C * p.N + 1) & ~(size_t)1) * 4) + 8 * 8;
  const size_t budget = 227 * 1024;
  // two resident CTAs per SM when three stages fit in half the shared memory
  int stages = 3, per_sm = 2;
  if (3 * stage_bytes + fixed > budget / 2) {
    per_sm = 1;
    stages = (int)((budget - fixed) / stage_bytes);
    if (stages > 4) stages = 4;
  }
  if (stages < 2) SVIT_FAIL(SVIT_ERR_UNSUPPORTED, "svit_aggregate: N=%d C=%d does not fit shared memory", p.N, p.C);
  p.stages = stages;
  const size_t smem = stages * stage_bytes + fixed;
  auto kern = aggregate_kernel<OutT, BLOCK>;
  SVIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t grid = (int64_t)sm_count() * per_sm;
  if (grid > p.num_tiles) grid = p.num_tiles;
  kern<<<(unsigned)grid, BLOCK, smem, stream>>>(p);
  SVIT_LAUNCH_CHECK("aggregate_kernel");
  return SVIT_OK;
}

template <typename OutT>
int dispatch_block(const AggParams& p, cudaStream_t stream) {
  if (p.N <= 16) return launch<OutT, 256>(p, stream);
  if (p.N <= 32) return launch<OutT, 128>(p, stream);
  return launch<OutT, 64>(p, stream);
}

}  // namespace
}  // namespace svit

extern "C" int svit_aggregate(const float* deltas, int64_t delta_stride, const float* w0, const float* ratios,
                              void* out, int64_t out_stride, int out_dtype, int64_t P, int N, int C,
                              svit_stream_t stream) {
  using namespace svit;
  SVIT_CHECK_ARG(deltas && ratios && out, "svit_aggregate: null pointer");
  SVIT_CHECK_ARG(P >= 0 && N >= 1 && N <= 64 && C >= 1 && C <= 256, "svit_aggregate: P=%lld N=%d C=%d out of range",
                 (long long)P, N, C);
  if (P == 0) return SVIT_OK;
  const int64_t p8 = round_up(P, 8);
  if (!aligned16(deltas) || !aligned16(out) || (w0 && !aligned16(w0)) || delta_stride % 8 || out_stride % 8 ||
      delta_stride < p8 || out_stride < p8)
    SVIT_FAIL(SVIT_ERR_ALIGN,
              "svit_aggregate: pointers must be 16-byte aligned and strides multiples of 8 and >= round_up(P, 8)");
  AggParams p{};
  p.deltas = deltas;
  p.delta_stride = delta_stride;
  p.w0 = w0;
  p.ratios = ratios;
  p.out = out;
  p.out_stride = out_stride;
  p.P = P;
  p.N = N;
  p.C = C;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (out_dtype) {
    case SVIT_F32: return dispatch_block<float>(p, s);
    case SVIT_BF16: return dispatch_block<__nv_bfloat16>(p, s);
    case SVIT_F16: return dispatch_block<__half>(p, s);
    default: SVIT_FAIL(SVIT_ERR_ARG, "svit_aggregate: unknown out_dtype %d", out_dtype);
  }
}
