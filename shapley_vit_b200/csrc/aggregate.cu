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

constexpr int kCChunk = 8;      // coalition accumulators held in registers per pass
constexpr int kVec = 8;         // consecutive parameters per thread
constexpr int kMaxRatios = 2048;  // C * N floats travelling as kernel parameters per launch

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
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity))
    if (clock64() - t0 > 4000000000LL) __trap();
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// 4 consecutive outputs, one streaming vector store (8 bytes for fp16/bf16, 16 for fp32)
template <typename OutT> struct Store4;
template <> struct Store4<float> {
  static __device__ __forceinline__ void st(float* p, const float* v) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
  }
};
template <> struct Store4<__nv_bfloat16> {
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float* v) {
    __stcs(reinterpret_cast<uint2*>(p), make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3])));
  }
};
template <> struct Store4<__half> {
  static __device__ __forceinline__ void st(__half* p, const float* v) {
    __stcs(reinterpret_cast<uint2*>(p), make_uint2(pack_f16x2_sat(v[0], v[1]), pack_f16x2_sat(v[2], v[3])));
  }
};

struct AggParams {
  const float* deltas;
  int64_t delta_stride;
  const float* w0;  // may be null
  void* out;
  int64_t out_stride;
  int64_t P;
  int N, C, stages;
  int64_t num_tiles;
  // FedAvg ratios [C, N], by value: they reach the SM through the constant bank, so the
  // membership test (ratio != 0) and the multiplier are warp-uniform operands, not loads.
  float ratios[kMaxRatios];
  // membership bit masks: masks[(c0 / kCChunk) * N + j] bit cc <=> ratios[(c0 + cc) * N + j] != 0
  uint32_t masks[kMaxRatios / kCChunk];
};

// dynamic smem: [STAGES][(N+1)][TILE] floats | mbarriers [STAGES]
template <typename OutT, int BLOCK>
__global__ void __launch_bounds__(BLOCK) aggregate_kernel(const __grid_constant__ AggParams p) {
  constexpr int TILE = BLOCK * kVec;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int N = p.N, C = p.C, S = p.stages;
  const int rows = N + 1;  // row N holds W_0
  float* stage_base = reinterpret_cast<float*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_base + (size_t)S * rows * TILE);
  const int tid = threadIdx.x;

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
    // Each thread owns two groups of 4 consecutive parameters, TILE/2 apart, so that every
    // 128-bit shared-memory read of a warp is contiguous (conflict-free).
    const int ea = tid * 4, eb = TILE / 2 + tid * 4;
    if (ea < len) {
      const bool has_b = eb < len;
      float w[kVec];
#pragma unroll
      for (int q = 0; q < kVec; ++q) w[q] = 0.f;
      if (p.w0) {
        const float4 a = *reinterpret_cast<const float4*>(st + (size_t)N * TILE + ea);
        w[0] = a.x, w[1] = a.y, w[2] = a.z, w[3] = a.w;
        if (has_b) {
          const float4 b = *reinterpret_cast<const float4*>(st + (size_t)N * TILE + eb);
          w[4] = b.x, w[5] = b.y, w[6] = b.z, w[7] = b.w;
        }
      }
      OutT* outp = reinterpret_cast<OutT*>(p.out) + t * TILE;
      for (int c0 = 0; c0 < C; c0 += kCChunk) {
        float acc[kCChunk][kVec];
#pragma unroll
        for (int cc = 0; cc < kCChunk; ++cc)
#pragma unroll
          for (int q = 0; q < kVec; ++q) acc[cc][q] = 0.f;
        for (int j = 0; j < N; ++j) {
          const float4 da = *reinterpret_cast<const float4*>(st + (size_t)j * TILE + ea);
          const float4 db = *reinterpret_cast<const float4*>(st + (size_t)j * TILE + eb);
          const float d[kVec] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
          // One mask word decides all 8 memberships; the 8 ratios are fetched unconditionally
          // (constant bank, warp-uniform) so no load sits on a branch's critical path.
          const uint32_t m = p.masks[(c0 / kCChunk) * N + j];
          float r[kCChunk];
#pragma unroll
          for (int cc = 0; cc < kCChunk; ++cc) {
            r[cc] = p.ratios[(c0 + cc) * N + j];
            asm volatile("" : "+f"(r[cc]));
          }
#pragma unroll
          for (int cc = 0; cc < kCChunk; ++cc) {
            if (m & (1u << cc)) {
#pragma unroll
              for (int q = 0; q < kVec; ++q) acc[cc][q] = __fadd_rn(acc[cc][q], __fmul_rn(r[cc], d[q]));
            }
          }
        }
#pragma unroll
        for (int cc = 0; cc < kCChunk; ++cc) {
          if (c0 + cc < C) {
            float v[kVec];
#pragma unroll
            for (int q = 0; q < kVec; ++q) v[q] = __fadd_rn(w[q], acc[cc][q]);
            OutT* o = outp + (size_t)(c0 + cc) * p.out_stride;
#pragma unroll
            for (int grp = 0; grp < 2; ++grp) {
              const int e = grp ? eb : ea;
              if (e + 4 <= len) {
                Store4<OutT>::st(o + e, v + 4 * grp);
              } else {  // ragged tail: element-wise
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  if (e + q < len) o[e + q] = Cvt<OutT>::from_f(v[4 * grp + q]);
              }
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
  constexpr int TILE = BLOCK * kVec;
  p.num_tiles = (p.P + TILE - 1) / TILE;
  const size_t stage_bytes = (size_t)(p.N + 1) * TILE * 4;
  const size_t fixed = 8 * 8;
  const size_t budget = 227 * 1024;
  int stages = 3;
  if (3 * stage_bytes + fixed > budget) stages = 2;
  if (stages * stage_bytes + fixed > budget)
    SVIT_FAIL(SVIT_ERR_UNSUPPORTED, "svit_aggregate: N=%d does not fit shared memory", p.N);
  p.stages = stages;
  const size_t smem = stages * stage_bytes + fixed;
  int per_sm = (int)(budget / (smem + 1024));  // +1 KB per-CTA reservation
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
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
  if (p.N <= 17) return launch<OutT, 128>(p, stream);  // stage = (N+1) * 4 KB
  if (p.N <= 35) return launch<OutT, 64>(p, stream);
  return launch<OutT, 32>(p, stream);
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
  SVIT_CHECK_ARG(out_dtype == SVIT_F32 || out_dtype == SVIT_BF16 || out_dtype == SVIT_F16,
                 "svit_aggregate: unknown out_dtype %d", out_dtype);
  if (P == 0) return SVIT_OK;
  const int64_t p8 = round_up(P, 8);
  if (!aligned16(deltas) || !aligned16(out) || (w0 && !aligned16(w0)) || delta_stride % 8 || out_stride % 8 ||
      delta_stride < p8 || out_stride < p8)
    SVIT_FAIL(SVIT_ERR_ALIGN,
              "svit_aggregate: pointers must be 16-byte aligned and strides multiples of 8 and >= round_up(P, 8)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int es = dtype_size(out_dtype);
  // the ratio rows travel as kernel parameters: at most kMaxRatios floats per launch
  const int c_per_launch = (kMaxRatios / N) / kCChunk * kCChunk;
  for (int c0 = 0; c0 < C; c0 += c_per_launch) {
    const int cn = C - c0 < c_per_launch ? C - c0 : c_per_launch;
    AggParams p{};
    p.deltas = deltas;
    p.delta_stride = delta_stride;
    p.w0 = w0;
    p.out = static_cast<char*>(out) + (size_t)c0 * out_stride * es;
    p.out_stride = out_stride;
    p.P = P;
    p.N = N;
    p.C = cn;
    // (p is value-initialised: ratios of the padding coalitions up to a multiple of 8 are 0)
    for (int i = 0; i < cn * N; ++i) p.ratios[i] = ratios[(size_t)c0 * N + i];
    for (int c = 0; c < cn; ++c)
      for (int j = 0; j < N; ++j)
        if (p.ratios[c * N + j] != 0.f) p.masks[(c / kCChunk) * N + j] |= 1u << (c % kCChunk);
    int rc;
    switch (out_dtype) {
      case SVIT_F32: rc = dispatch_block<float>(p, s); break;
      case SVIT_BF16: rc = dispatch_block<__nv_bfloat16>(p, s); break;
      default: rc = dispatch_block<__half>(p, s); break;
    }
    if (rc) return rc;
  }
  return SVIT_OK;
}
