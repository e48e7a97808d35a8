// K1 -- coalition aggregation (HBM-bound).
//
// Replaces, for a whole batch of C coalitions at once, the reference's
//   get_aggregated_model      federated_learning/utils.py:781-792   agg = sum_j r_j * Delta_j
//   ServerBase.model_agg_lazy federated_learning/server2.py:121-127 W_S = W_0 + agg
// which the reference runs once per coalition and per state_dict key (2|S|-1 elementwise
// launches per key, each re-reading its operands).
//
// Data movement: the stacked deltas [N, P] and W_0 [P] are streamed from HBM exactly once.
// A persistent CTA owns tiles of 1 024 consecutive parameters; one elected thread stages the rows
// of a tile into shared memory, <= 8 client rows (+ W_0) per transaction set, with 1-D TMA bulk
// copies (cp.async.bulk, completion on an mbarrier), two stages deep, so the loads in flight do
// not depend on occupancy or registers and shared memory per CTA does not grow with N.  Each
// thread owns 2 x 4 consecutive parameters, keeps 8 coalition accumulators (x8 lanes) in
// registers, walks the clients in ascending order reading its float4 from shared memory, and
// writes every coalition's row with 64/128-bit streaming stores.  fp32 outputs: FMUL2 (packed
// fp32x2 products) + FADD, three issue slots per two parameters, the reference's two roundings;
// 16-bit outputs: one FFMA2 (scalar ratio x packed pair), accumulators started from W_0.
//
// Arithmetic (fp32 output): every product and every sum is a separately rounded fp32 operation
// (__fmul_rn/__fadd_rn; ptxas would otherwise contract to FMA), in ascending client order,
// which is bit-identical to the reference whenever frozenset(coalition) iterates in ascending
// order.  Algorithmic bytes per launch:
//   4*P*(N+1) + sizeof(out)*P*C.
#include <cstdlib>
#include <type_traits>

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

// r * d for two parameters in one issue slot (sm_100 FMUL2, IEEE round-to-nearest like FMUL).
// The sums stay scalar FADDs on purpose: ptxas 12.9 contracts a packed product feeding a packed
// sum into one FFMA2 even when both carry .rn (it does not for scalar code), and a fused
// multiply-add rounds once where the reference (`ratio * delta`, then `agg + ...`, on fp32
// tensors) rounds twice.
__device__ __forceinline__ float2 mul2(float r, float2 d) {
  float2 o;
  asm("{.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %2};\n\tmov.b64 rb, {%3, %4};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;}"
      : "=f"(o.x), "=f"(o.y)
      : "f"(r), "f"(d.x), "f"(d.y));
  return o;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
// acc + r * d with ONE rounding (FFMA2): only for 16-bit outputs, see accumulate()
__device__ __forceinline__ float2 fma2(float r, float2 d, float2 acc) {
  float2 o;
  asm("{.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %2};\n\tmov.b64 rb, {%3, %4};\n\tmov.b64 rc, {%5, %6};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;}"
      : "=f"(o.x), "=f"(o.y)
      : "f"(r), "f"(d.x), "f"(d.y), "f"(acc.x), "f"(acc.y));
  return o;
}
// fp32 outputs reproduce the reference's arithmetic exactly (product rounded, then the sum).  16-bit
// outputs -- the operand feed of the 16-bit GEMMs, which the reference has no counterpart of -- use a
// fused multiply-add: the result is rounded to 11 / 8 mantissa bits right afterwards, the fused
// operation is at most one fp32 ulp away from the two-rounding value before that, and it takes a third
// of the issue slots, which is what keeps the kernel HBM-bound at power-capped SM clocks.
// Output element types.  OutX3 / OutC8 are the split operand formats (packed operand arrays, include/svit.h).
// OutX3 (hi + lo fp16, ~21 bits) is split from the reference's two-rounding fp32 value, like the fp32 output;
// OutC8 (fp16 + e4m3 residual, ~16 bits) from the fused accumulation, like the 16-bit outputs: one fp32 ulp is
// 2^-8 of what the format resolves, and the fused fold is what keeps the kernel on the HBM roofline at the
// power-capped clock of the bench step (0.83 -> of the copy peak with the two-rounding fold).
struct OutX3 { uint32_t tag; };
struct OutC8 { uint32_t tag; };
template <typename OutT> struct Fused { static constexpr bool value = sizeof(OutT) != 4; };
template <> struct Fused<OutC8> { static constexpr bool value = true; };

template <typename OutT>
__device__ __forceinline__ float2 accumulate(float2 acc, float r, float2 d) {
  if constexpr (!Fused<OutT>::value)
    return add2(acc, mul2(r, d));
  else
    return fma2(r, d, acc);
}

// 4 consecutive outputs at element index i of the output array: streaming vector stores
// (8 bytes for fp16/bf16, 16 for fp32; one store per plane for the split formats)
template <typename OutT> struct Store4;
template <> struct Store4<float> {
  static __device__ __forceinline__ void st(void* out, int64_t, int64_t i, float2 a, float2 b) {
    __stcs(reinterpret_cast<float4*>(static_cast<float*>(out) + i), make_float4(a.x, a.y, b.x, b.y));
  }
};
template <> struct Store4<__nv_bfloat16> {
  static __device__ __forceinline__ void st(void* out, int64_t, int64_t i, float2 a, float2 b) {
    __stcs(reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(out) + i), make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(b.x, b.y)));
  }
};
template <> struct Store4<__half> {
  static __device__ __forceinline__ void st(void* out, int64_t, int64_t i, float2 a, float2 b) {
    __stcs(reinterpret_cast<uint2*>(static_cast<__half*>(out) + i), make_uint2(pack_f16x2_sat(a.x, a.y), pack_f16x2_sat(b.x, b.y)));
  }
};
template <> struct Store4<OutX3> {
  static __device__ __forceinline__ void st(void* out, int64_t alloc, int64_t i, float2 a, float2 b) {
    uint32_t h0, h1, l0, l1;
    split_x3(a.x, a.y, h0, l0);
    split_x3(b.x, b.y, h1, l1);
    __stcs(reinterpret_cast<uint2*>(static_cast<__half*>(out) + i), make_uint2(h0, h1));
    __stcs(reinterpret_cast<uint2*>(reinterpret_cast<__half*>(static_cast<char*>(out) + 2 * alloc) + i), make_uint2(l0, l1));
  }
};
template <> struct Store4<OutC8> {
  static __device__ __forceinline__ void st(void* out, int64_t alloc, int64_t i, float2 a, float2 b) {
    uint32_t h0, h1;
    uint16_t a0, a1, b0, b1;
    split_c8(a.x, a.y, h0, a0, b0);
    split_c8(b.x, b.y, h1, a1, b1);
    __stcs(reinterpret_cast<uint2*>(static_cast<__half*>(out) + i), make_uint2(h0, h1));
    __stcs(reinterpret_cast<uint32_t*>(static_cast<char*>(out) + 2 * alloc + i), (uint32_t)a0 | ((uint32_t)a1 << 16));
    __stcs(reinterpret_cast<uint32_t*>(static_cast<char*>(out) + 3 * alloc + i), (uint32_t)b0 | ((uint32_t)b1 << 16));
  }
};
// one output (ragged tails)
template <typename OutT>
__device__ __forceinline__ void store1(void* out, int64_t alloc, int64_t i, float v) {
  if constexpr (std::is_same<OutT, OutX3>::value) {
    uint32_t h, l;
    split_x3(v, 0.f, h, l);
    static_cast<uint16_t*>(out)[i] = (uint16_t)h;
    reinterpret_cast<uint16_t*>(static_cast<char*>(out) + 2 * alloc)[i] = (uint16_t)l;
  } else if constexpr (std::is_same<OutT, OutC8>::value) {
    uint32_t h;
    uint16_t a, b;
    split_c8(v, 0.f, h, a, b);
    static_cast<uint16_t*>(out)[i] = (uint16_t)h;
    reinterpret_cast<uint8_t*>(out)[2 * alloc + i] = (uint8_t)a;
    reinterpret_cast<uint8_t*>(out)[3 * alloc + i] = (uint8_t)b;
  } else {
    static_cast<OutT*>(out)[i] = Cvt<OutT>::from_f(v);
  }
}

struct AggParams {
  const float* deltas;
  int64_t delta_stride;
  const float* w0;  // shared base row [P], may be null
  const float* base;  // per-coalition base rows [C, base_stride] fp32 (svit_aggregate_onto), exclusive with w0
  int64_t base_stride;
  void* out;          // output array (split formats: the main plane of the packed operand array)
  int64_t out_off;    // element offset of this launch's first coalition row inside `out`
  int64_t out_alloc;  // plane pitch in elements (split formats)
  int64_t out_stride;
  int64_t P;
  int N, C, stages;
  int group;  // client rows staged per work item (8, or N when all rows of a multi-chunk launch fit: see launch())
  int64_t num_tiles;
  // FedAvg ratios, by value, in [chunk of 8 coalitions][client j][8] order (0 = not a member or
  // padding), and one membership word per (chunk, j): bit cc <=> coalition 8*chunk+cc contains j.
  // Each CTA copies both tables to shared memory once; the inner loop then reads 8 ratios and
  // the mask with three broadcast loads.
  float ratios[kMaxRatios];
  uint32_t masks[kMaxRatios / kCChunk];
};


// base[c] values of one thread's two parameter groups (per-coalition base rows of svit_aggregate_onto)
__device__ __forceinline__ void load_base(const float* row, int ea, int eb, int len, float2 (&w)[4]) {
  float f[8];
#pragma unroll
  for (int grp = 0; grp < 2; ++grp) {
    const int e = grp ? eb : ea;
    if (e + 4 <= len) {
      const float4 a = *reinterpret_cast<const float4*>(row + e);
      f[4 * grp] = a.x, f[4 * grp + 1] = a.y, f[4 * grp + 2] = a.z, f[4 * grp + 3] = a.w;
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) f[4 * grp + q] = e + q < len ? row[e + q] : 0.f;
    }
  }
  w[0] = make_float2(f[0], f[1]), w[1] = make_float2(f[2], f[3]);
  w[2] = make_float2(f[4], f[5]), w[3] = make_float2(f[6], f[7]);
}

constexpr int kBlock = 128;
constexpr int kTile = kBlock * kVec;  // 1024 parameters = 4 KB per staged row
constexpr int kGroup = 8;             // client rows staged per work item (default)
constexpr int kThreads = kBlock + 32;  // four consumer warps + the producer warp

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// the 128 consumer threads only (the producer warp never joins): ragged-tile staging
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// One kernel for every (N, C).  A persistent CTA walks its tiles; the work of a tile is split into
// ITEMS, each one TMA transaction set of <= 8 client rows (+ the W_0 row with a tile's first item):
//   N <= 8 : one item per tile; every chunk of 8 coalitions folds from the same staged rows;
//   N  > 8 : (chunk, group of 8 clients) items, chunk-major -- the accumulators of a chunk stay in
//            registers across its groups, the rows are fetched again per chunk (L2 hits: the same
//            CTA read them microseconds earlier).
// Shared memory per CTA therefore does not grow with N (2 stages x 9 rows x 4 KB = 72 KB, three
// CTAs per SM for any N), the barrier traffic is one wait per 8 rows, and the inner loop is the
// same two-way unrolled, software-pipelined fold for every shape.
// dynamic smem: [S][rows_per_stage][kTile] floats | ratio table | mask table | mbarriers [S]
// MULTI = several items per tile (N > rows per item); !MULTI lets the compiler see that g == 0 always,
// i.e. that the accumulators die at the end of every item.
template <typename OutT, bool MULTI>
__global__ void __launch_bounds__(kThreads, 3) aggregate_kernel(const __grid_constant__ AggParams p) {
  constexpr int TILE = kTile;
  constexpr int ROW4 = TILE / 4;  // float4 per staged row
  // 16-bit outputs start their accumulators from W_0 (one rounding fewer per output and no add in the
  // epilogue; the value is rounded to 11 / 8 bits right after); fp32 outputs keep the reference's
  // order W_0 + (sum_j r_j d_j), so their accumulators start from zero
  constexpr bool kFromW0 = Fused<OutT>::value;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int N = p.N, C = p.C, S = p.stages;
  const int nchunks = (C + kCChunk - 1) / kCChunk;
  const int gsz = p.group;
  const int G = MULTI ? (N + gsz - 1) / gsz : 1;
  const int srows = min(N, gsz) + 1;         // slot srows - 1 holds the W_0 row
  const int ipt = G == 1 ? 1 : nchunks * G;  // items per tile
  float* stage_base = reinterpret_cast<float*>(smem_raw);
  float* s_ratio = stage_base + (size_t)S * srows * TILE;
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(s_ratio + (size_t)nchunks * N * kCChunk);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_mask + ((nchunks * N + 1) & ~1));  // full[S]
  uint64_t* empty = bars + S;
  const int tid = threadIdx.x;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&bars[s], 1);
      mbar_init(&empty[s], kBlock / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < nchunks * N * kCChunk; i += kThreads) s_ratio[i] = p.ratios[i];
  for (int i = tid; i < nchunks * N; i += kThreads) s_mask[i] = p.masks[i];
  __syncthreads();

  const int64_t P = p.P;
  auto tile_len = [&](int64_t t) -> int { return (int)min((int64_t)TILE, P - t * TILE); };
  // a tile goes through TMA when its length is a multiple of 4 floats (16 B granules); the ragged
  // last tile (at most one per launch) is staged by all threads in the consumer path
  auto issue = [&](int64_t t, int sub, int s) {  // thread 0 only
    const int len = tile_len(t);
    if (len & 3) return;
    const int g = MULTI ? sub % G : 0, j0 = g * gsz, nj = min(gsz, N - j0);
    const bool with_w0 = g == 0 && p.w0 != nullptr;
    float* dst = stage_base + (size_t)s * srows * TILE;
    const uint32_t bytes = (uint32_t)len * 4u;
    mbar_expect_tx(&bars[s], bytes * (uint32_t)(nj + (with_w0 ? 1 : 0)));
    for (int j = 0; j < nj; ++j)
      bulk_g2s(dst + (size_t)j * TILE, p.deltas + (size_t)(j0 + j) * p.delta_stride + t * TILE, bytes, &bars[s]);
    if (with_w0) bulk_g2s(dst + (size_t)(srows - 1) * TILE, p.w0 + t * TILE, bytes, &bars[s]);
  };
  // ===== producer warp: walks the same item sequence as the consumers, one stage ahead of their releases =====
  if (tid >= kBlock) {
    if (tid == kBlock) {
      int ps_ = 0;
      uint32_t pph = 0;
      for (int64_t t = blockIdx.x; t < p.num_tiles; t += gridDim.x)
        for (int sub = 0; sub < ipt; ++sub) {
          mbar_wait(&empty[ps_], pph ^ 1);  // all four consumer warps have released the stage (free at start)
          issue(t, sub, ps_);
          if (++ps_ == S) ps_ = 0, pph ^= 1;
        }
    }
    return;
  }

  // Each thread owns two groups of 4 consecutive parameters, TILE/2 apart, so that every 128-bit
  // shared-memory read of a warp is contiguous (conflict-free).
  const int ea = tid * 4, eb = TILE / 2 + tid * 4;
  float2 acc[kCChunk][4];
  float2 w[4];
  int s = 0;
  uint32_t parity = 0;
  for (int64_t t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
    const int len = tile_len(t);
    const int64_t outp = p.out_off + t * TILE;  // element index of the tile in coalition row 0

    // the chunk's coalition rows: W_0 (or the per-coalition base) + acc, converted and stored
    auto write_chunk = [&](int ch) {
      if (len == TILE && !p.base) {  // full tile, shared W_0: whole vectors, the row pointer just advances
        int64_t o = outp + (int64_t)ch * kCChunk * p.out_stride;
        const int ncc = min(kCChunk, C - ch * kCChunk);
#pragma unroll
        for (int cc = 0; cc < kCChunk; ++cc) {
          if (cc < ncc) {
            float2 v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = kFromW0 ? acc[cc][q] : add2(w[q], acc[cc][q]);
            Store4<OutT>::st(p.out, p.out_alloc, o + ea, v[0], v[1]);
            Store4<OutT>::st(p.out, p.out_alloc, o + eb, v[2], v[3]);
            o += p.out_stride;
          }
        }
        return;
      }
#pragma unroll
      for (int cc = 0; cc < kCChunk; ++cc) {
        const int c = ch * kCChunk + cc;
        if (c < C) {
          float2 v[4], wb[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) wb[q] = w[q];
          if (p.base) load_base(p.base + (size_t)c * p.base_stride + t * TILE, ea, eb, len, wb);
#pragma unroll
          for (int q = 0; q < 4; ++q) v[q] = (kFromW0 && !p.base) ? acc[cc][q] : add2(wb[q], acc[cc][q]);
          const int64_t o = outp + (int64_t)c * p.out_stride;
#pragma unroll
          for (int grp = 0; grp < 2; ++grp) {
            const int e = grp ? eb : ea;
            if (e + 4 <= len) {
              Store4<OutT>::st(p.out, p.out_alloc, o + e, v[2 * grp], v[2 * grp + 1]);
            } else {  // ragged tail: element-wise
              const float f[4] = {v[2 * grp].x, v[2 * grp].y, v[2 * grp + 1].x, v[2 * grp + 1].y};
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (e + q < len) store1<OutT>(p.out, p.out_alloc, o + e + q, f[q]);
            }
          }
        }
      }
    };

    for (int sub = 0; sub < ipt; ++sub) {
      const int g = MULTI ? sub % G : 0, ch0 = MULTI ? sub / G : 0;  // (G == 1: the chunk loop runs below)
      const int j0 = g * gsz, nj = min(gsz, N - j0);
      float* st = stage_base + (size_t)s * srows * TILE;
      if (len & 3) {  // ragged tail: guarded scalar staging of this item's rows
        consumer_sync();  // nobody still reads an earlier item from this stage
        for (int j = 0; j <= nj; ++j) {
          const float* src = j < nj ? p.deltas + (size_t)(j0 + j) * p.delta_stride + t * TILE
                                    : ((g == 0 && p.w0) ? p.w0 + t * TILE : nullptr);
          float* dst = st + (size_t)(j < nj ? j : srows - 1) * TILE;
          if (j == nj && g != 0) break;
          for (int e = tid; e < TILE; e += kBlock) dst[e] = (src && e < len) ? src[e] : 0.f;
        }
        consumer_sync();
      } else {
        mbar_wait(&bars[s], parity);
      }
      if (ea < len) {
        if (g == 0) {  // first item of a tile (G == 1) or of a (tile, chunk): W_0 arrives with it
#pragma unroll
          for (int q = 0; q < 4; ++q) w[q] = make_float2(0.f, 0.f);
          if (p.w0) {
            const float4 a = *reinterpret_cast<const float4*>(st + (size_t)(srows - 1) * TILE + ea);
            const float4 b = *reinterpret_cast<const float4*>(st + (size_t)(srows - 1) * TILE + eb);
            w[0] = make_float2(a.x, a.y), w[1] = make_float2(a.z, a.w);
            w[2] = make_float2(b.x, b.y), w[3] = make_float2(b.z, b.w);
          }
        }
        const int nch = G == 1 ? nchunks : 1;  // chunks folded from this item's rows
#pragma unroll 1
        for (int k = 0; k < nch; ++k) {
          const int ch = G == 1 ? k : ch0;
          if (g == 0) {
#pragma unroll
            for (int cc = 0; cc < kCChunk; ++cc)
#pragma unroll
              for (int q = 0; q < 4; ++q) acc[cc][q] = (kFromW0 && !p.base) ? w[q] : make_float2(0.f, 0.f);
          }
          const float4* rt = reinterpret_cast<const float4*>(s_ratio + ((size_t)ch * N + j0) * kCChunk);
          const uint32_t* mk = s_mask + ch * N + j0;
          // one client's contribution to the chunk's 8 coalitions
          auto fold = [&](const float4& da, const float4& db, uint32_t mc, const float4& r0, const float4& r1) {
            const float2 d[4] = {make_float2(da.x, da.y), make_float2(da.z, da.w), make_float2(db.x, db.y),
                                 make_float2(db.z, db.w)};
            const float r[kCChunk] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
            for (int cc = 0; cc < kCChunk; ++cc) {
              if (mc & (1u << cc)) {  // warp-uniform
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[cc][q] = accumulate<OutT>(acc[cc][q], r[cc], d[q]);
              }
            }
          };
          const float4* pa = reinterpret_cast<const float4*>(st + ea);
          const float4* pb = reinterpret_cast<const float4*>(st + eb);
          // two-deep software pipeline, unrolled by two clients so the prefetched registers are used where
          // they land (no rotation moves): the loads of client j + 1 are in flight during the arithmetic of j
          float4 da0 = pa[0], db0 = pb[0], r00 = rt[0], r01 = rt[1];
          uint32_t m0 = mk[0];
          int rem = nj;  // clients left, counting the one already in set 0 (compared with immediates only)
#pragma unroll 1
          for (; rem >= 2; rem -= 2) {
            const float4 da1 = pa[ROW4], db1 = pb[ROW4], r10 = rt[2], r11 = rt[3];
            const uint32_t m1 = mk[1];
            fold(da0, db0, m0, r00, r01);
            pa += 2 * ROW4, pb += 2 * ROW4, rt += 4, mk += 2;
            if (rem > 2) da0 = pa[0], db0 = pb[0], r00 = rt[0], r01 = rt[1], m0 = mk[0];
            fold(da1, db1, m1, r10, r11);
          }
          if (rem == 1) fold(da0, db0, m0, r00, r01);
          if (g == G - 1) write_chunk(ch);
        }
      }
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&empty[s]);  // this warp is done reading stage s
      if (++s == S) {
        s = 0;
        parity ^= 1;
      }
    }
  }
}

template <typename OutT>
int launch(const AggParams& base, cudaStream_t stream) {
  AggParams p = base;
  p.num_tiles = (p.P + kTile - 1) / kTile;
  const int nchunks = (p.C + kCChunk - 1) / kCChunk;
  // Rows per item: 8.  With several chunks of coalitions AND more than 8 clients the rows would be
  // fetched once per chunk; up to N = 13 the whole tile (N + 1 rows, two stages, two CTAs per SM) still
  // fits, every chunk folds from the same staged rows and nothing is fetched twice.
  p.group = (nchunks > 1 && p.N > kGroup && p.N <= 13) ? p.N : kGroup;
  const int srows = (p.N < p.group ? p.N : p.group) + 1;
  const size_t stage_bytes = (size_t)srows * kTile * 4;
  const size_t fixed = (size_t)nchunks * p.N * kCChunk * 4 + (size_t)((nchunks * p.N + 1) & ~1) * 4 + 2 * 8 * 8;
  const size_t budget = 227 * 1024;
  // Two stages per CTA and as many CTAs per SM as fit: the loads in flight per SM are the same as
  // with deeper rings, but more warps hide the shared-memory latency of the inner loop.
  int stages = 2;
  if (4 * stage_bytes + fixed + 1024 <= budget / 3) stages = 4;
  else if (3 * stage_bytes + fixed + 1024 <= budget / 3) stages = 3;
  p.stages = stages;
  const size_t smem = stages * stage_bytes + fixed;
  if (smem > budget) SVIT_FAIL(SVIT_ERR_UNSUPPORTED, "svit_aggregate: ratio tables of N=%d, C=%d do not fit shared memory", p.N, p.C);
  auto kern = p.N > p.group ? aggregate_kernel<OutT, true> : aggregate_kernel<OutT, false>;
  SVIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;  // resident CTAs per SM (registers and shared memory): the grid is exactly one wave
  SVIT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
  if (per_sm < 1) per_sm = 1;
  int64_t grid = (int64_t)sm_count() * per_sm;
  if (grid > p.num_tiles) grid = p.num_tiles;
  kern<<<(unsigned)grid, kThreads, smem, stream>>>(p);
  SVIT_LAUNCH_CHECK("aggregate_kernel");
  return SVIT_OK;
}

template <typename OutT>
int dispatch_block(const AggParams& p, cudaStream_t stream) {
  // (a variant with 4 or 2 parameters and 16 or 32 coalitions per thread -- half / a quarter of the passes
  // over a tile's rows when N > 8 and C > 8 -- was measured slower everywhere once the producer warp had
  // removed the per-item block barrier: 0.45 vs 0.60 of the copy peak at N = 16, C = 32)
  return launch<OutT>(p, stream);
}

}  // namespace
}  // namespace svit

namespace svit {
namespace {

int aggregate_impl(const char* who, const float* deltas, int64_t delta_stride, const float* w0, const float* base,
                   int64_t base_stride, const float* ratios, void* out, int64_t out_off, int64_t out_stride, int out_dtype,
                   int out_fmt, int64_t out_alloc, int64_t P, int N, int C, svit_stream_t stream) {
  SVIT_CHECK_ARG(deltas && ratios && out, "%s: null pointer", who);
  SVIT_CHECK_ARG(P >= 0 && N >= 1 && N <= 64 && C >= 1 && C <= 256, "%s: P=%lld N=%d C=%d out of range", who, (long long)P, N,
                 C);
  SVIT_CHECK_ARG(out_dtype == SVIT_F32 || out_dtype == SVIT_BF16 || out_dtype == SVIT_F16, "%s: unknown out_dtype %d", who,
                 out_dtype);
  SVIT_CHECK_ARG(out_fmt == SVIT_FMT_PLAIN || ((out_fmt == SVIT_FMT_X3 || out_fmt == SVIT_FMT_C8) && out_alloc % 16 == 0 &&
                                               out_off % 8 == 0 && out_off >= 0 &&
                                               out_alloc >= out_off + (int64_t)(C - 1) * out_stride + P),
                 "%s: bad output format %d / plane pitch %lld", who, out_fmt, (long long)out_alloc);
  if (P == 0) return SVIT_OK;
  const int64_t p8 = round_up(P, 8);
  if (!aligned16(deltas) || !aligned16(out) || (w0 && !aligned16(w0)) || (base && !aligned16(base)) || delta_stride % 8 ||
      out_stride % 8 || delta_stride < p8 || out_stride < p8 || (base && (base_stride % 8 || base_stride < p8)))
    SVIT_FAIL(SVIT_ERR_ALIGN, "%s: pointers must be 16-byte aligned and strides multiples of 8 and >= round_up(P, 8)", who);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // the ratio rows travel as kernel parameters: at most kMaxRatios floats per launch
  const int c_per_launch = (kMaxRatios / N) / kCChunk * kCChunk;
  for (int c0 = 0; c0 < C; c0 += c_per_launch) {
    const int cn = C - c0 < c_per_launch ? C - c0 : c_per_launch;
    AggParams p{};
    p.deltas = deltas;
    p.delta_stride = delta_stride;
    p.w0 = w0;
    p.base = base ? base + (size_t)c0 * base_stride : nullptr;
    p.base_stride = base_stride;
    p.out = out;
    p.out_off = out_off + (int64_t)c0 * out_stride;
    p.out_alloc = out_alloc;
    p.out_stride = out_stride;
    p.P = P;
    p.N = N;
    p.C = cn;
    // p is value-initialised: ratios of the padding coalitions up to a multiple of 8 stay 0
    for (int c = 0; c < cn; ++c)
      for (int j = 0; j < N; ++j) {
        const float r = ratios[(size_t)(c0 + c) * N + j];
        p.ratios[((size_t)(c / kCChunk) * N + j) * kCChunk + c % kCChunk] = r;
        if (r != 0.f) p.masks[(c / kCChunk) * N + j] |= 1u << (c % kCChunk);
      }
    int rc;
    if (out_fmt == SVIT_FMT_X3) rc = dispatch_block<OutX3>(p, s);
    else if (out_fmt == SVIT_FMT_C8) rc = dispatch_block<OutC8>(p, s);
    else
      switch (out_dtype) {
        case SVIT_F32: rc = dispatch_block<float>(p, s); break;
        case SVIT_BF16: rc = dispatch_block<__nv_bfloat16>(p, s); break;
        default: rc = dispatch_block<__half>(p, s); break;
      }
    if (rc) return rc;
  }
  return SVIT_OK;
}

}  // namespace
}  // namespace svit

extern "C" int svit_aggregate(const float* deltas, int64_t delta_stride, const float* w0, const float* ratios,
                              void* out, int64_t out_stride, int out_dtype, int64_t P, int N, int C,
                              svit_stream_t stream) {
  return svit::aggregate_impl("svit_aggregate", deltas, delta_stride, w0, nullptr, 0, ratios, out, 0, out_stride, out_dtype,
                              SVIT_FMT_PLAIN, 0, P, N, C, stream);
}

extern "C" int svit_aggregate_onto(const float* deltas, int64_t delta_stride, const float* base, int64_t base_stride,
                                   const float* ratios, void* out, int64_t out_stride, int out_dtype, int64_t P, int N,
                                   int C, svit_stream_t stream) {
  using namespace svit;
  SVIT_CHECK_ARG(base != nullptr, "svit_aggregate_onto: base is null");
  return aggregate_impl("svit_aggregate_onto", deltas, delta_stride, nullptr, base, base_stride, ratios, out, 0, out_stride,
                        out_dtype, SVIT_FMT_PLAIN, 0, P, N, C, stream);
}

extern "C" int svit_aggregate_split(const float* deltas, int64_t delta_stride, const float* w0, const float* base,
                                    int64_t base_stride, const float* ratios, void* out, int64_t out_off,
                                    int64_t out_stride, int64_t out_alloc, int out_fmt, int64_t P, int N, int C,
                                    svit_stream_t stream) {
  using namespace svit;
  SVIT_CHECK_ARG(out_fmt == SVIT_FMT_X3 || out_fmt == SVIT_FMT_C8, "svit_aggregate_split: out_fmt must be a split format");
  SVIT_CHECK_ARG(!(w0 && base), "svit_aggregate_split: give w0 (shared) or base (per coalition), not both");
  return aggregate_impl("svit_aggregate_split", deltas, delta_stride, w0, base, base_stride, ratios, out, out_off, out_stride,
                        SVIT_F16, out_fmt, out_alloc, P, N, C, stream);
}
