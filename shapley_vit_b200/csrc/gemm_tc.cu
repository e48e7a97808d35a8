// K2 -- grouped GEMM on the 5th-generation tensor cores (tcgen05 + TMEM accumulators, TMA
// operand staging), hand-written PTX for sm_100a.
//
// Serves every dense contraction of the batched ViT forward the reference runs through
// HF/cuBLAS one coalition at a time (federated_learning/utils.py:886; HF modeling_vit.py
// :228-230, :265-268, :287-312, patch embedding :166):
//     out[g] = epilogue( A[g] (M x K) * B[g]^T (N x K) ),   g = coalition,
// both operands K-contiguous ("K-major"), fp32 accumulation in tensor memory.
//
// Structure: persistent, warp-specialised CTAs, one per SM (launch_bounds(320, 1)):
//   warp 0      TMA producer   cp.async.bulk.tensor.3d (k, row, group) -> 128B-swizzled smem ring
//   warp 1      MMA issuer     one thread issues tcgen05.mma (128 x BN x 16|8 per instruction),
//                              tcgen05.commit releases smem stages / publishes accumulators;
//                              also owns tcgen05.alloc / dealloc
//   warps 2..9  epilogue       tcgen05.ld 32x32b from TMEM -> +bias / GELU / +pos / +residual ->
//                              vector stores.  Two TMEM accumulator buffers, so the epilogue of
//                              tile i overlaps the MMAs of tile i+1.
// Tile order: n fastest, then m, then group, so CTAs running concurrently share A rows in L2.
// Operand precisions: bf16 / fp16 (kind::f16, K=16 per MMA) and tf32 (kind::tf32, K=8).
#include <cuda.h>

#include "epilogue.cuh"

namespace svit {
namespace {

constexpr int BM = 128;
constexpr int kEpiWarps = 8;
constexpr int kThreads = (2 + kEpiWarps) * 32;

template <int BN>
struct Cfg {
  static constexpr int STAGES = BN == 256 ? 4 : 6;
  static constexpr int A_BYTES = BM * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int NUM_BARS = 2 * STAGES + 4;
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + NUM_BARS * 8 + 16 + 1024;
};

// ---- PTX wrappers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// Bounded wait: a broken pipeline traps (launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
template <int KIND>
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  if constexpr (KIND == 0) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
  }
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (row) i
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

struct TcShape {
  int G, M, N, K;
  int tiles_m, tiles_n, num_kb;
  int64_t total_tiles;
  int a_grouped;  // 0: A shared by all groups
};

// ---- epilogue for 32 consecutive columns of one output row ---------------------------------
__device__ __forceinline__ void epilogue_row32(const EpiArgs& e, int g, int r, int n, float* v) {
  const int64_t orow = epi_out_row(e, r);
  const int N = e.N;
  if (n + 32 <= N && (N & 7) == 0) {
    if (e.bias) {
      const float4* b = reinterpret_cast<const float4*>(e.bias + (size_t)g * e.bias_gs + n);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 t = __ldg(b + i);
        v[4 * i] += t.x, v[4 * i + 1] += t.y, v[4 * i + 2] += t.z, v[4 * i + 3] += t.w;
      }
    }
    if (e.gelu) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
    }
    if (e.rowvec) {
      const float4* b = reinterpret_cast<const float4*>(e.rowvec + (size_t)g * e.rowvec_gs + (size_t)(orow % e.rows_out) * N + n);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 t = __ldg(b + i);
        v[4 * i] += t.x, v[4 * i + 1] += t.y, v[4 * i + 2] += t.z, v[4 * i + 3] += t.w;
      }
    }
    if (e.residual) {
      const float4* b = reinterpret_cast<const float4*>(e.residual + (size_t)g * e.residual_gs + (size_t)orow * N + n);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 t = b[i];
        v[4 * i] += t.x, v[4 * i + 1] += t.y, v[4 * i + 2] += t.z, v[4 * i + 3] += t.w;
      }
    }
    const size_t idx = (size_t)g * e.out_gs + (size_t)orow * N + n;
    if (e.out_dtype == SVIT_F32) {
      float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + idx);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else if (e.out_dtype == SVIT_BF16) {
      uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(e.out) + idx);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        o[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                          pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
    } else {
      uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(e.out) + idx);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        o[i] = make_uint4(pack_f16x2_sat(v[8 * i], v[8 * i + 1]), pack_f16x2_sat(v[8 * i + 2], v[8 * i + 3]),
                          pack_f16x2_sat(v[8 * i + 4], v[8 * i + 5]), pack_f16x2_sat(v[8 * i + 6], v[8 * i + 7]));
    }
  } else {  // ragged N: element-wise (static indices keep v[] in registers)
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (n + i < N) epi_store(e, g, orow, n + i, epi_apply(e, g, orow, n + i, v[i]));
  }
}

// ---- the kernel --------------------------------------------------------------------------
template <int BN, int KIND>
__global__ void __launch_bounds__(kThreads, 1)
    gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                   const TcShape sh, const EpiArgs epi, const uint32_t idesc) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)C::STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tfull_bar = empty_bar + C::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_b) : "memory");
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], kEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t tiles_per_group = (int64_t)sh.tiles_m * sh.tiles_n;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      int s = 0;
      uint32_t ph = 0;
      for (int64_t tile = blockIdx.x; tile < sh.total_tiles; tile += gridDim.x) {
        const int g = (int)(tile / tiles_per_group);
        const int rem = (int)(tile % tiles_per_group);
        const int m0 = (rem / sh.tiles_n) * BM, n0 = (rem % sh.tiles_n) * BN;
        for (int kb = 0; kb < sh.num_kb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + (size_t)s * C::STAGE_BYTES;
          mbar_expect_tx(&full_bar[s], C::STAGE_BYTES);
          const int k0 = kb * (KIND == 0 ? 64 : 32);
          tma_load_3d(sa, &tma_a, &full_bar[s], k0, m0, sh.a_grouped ? g : 0);
          tma_load_3d(sa + C::A_BYTES, &tma_b, &full_bar[s], k0, n0, g);
          if (++s == C::STAGES) s = 0, ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      int s = 0, acc = 0;
      uint32_t ph = 0, aph = 0;
      for (int64_t tile = blockIdx.x; tile < sh.total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < sh.num_kb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + (size_t)s * C::STAGE_BYTES);
          const uint32_t b_addr = a_addr + C::A_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k)  // 4 x 32 bytes along K inside the 128-byte swizzle span
            tc_mma<KIND>(d_tmem, umma_desc(a_addr + k * 32), umma_desc(b_addr + k * 32), idesc, (uint32_t)(kb | k));
          tc_commit(&empty_bar[s]);  // smem stage reusable once these MMAs have read it
          if (++s == C::STAGES) s = 0, ph ^= 1;
        }
        tc_commit(&tfull_bar[acc]);  // accumulator complete
        if ((acc ^= 1) == 0) aph ^= 1;
      }
    }
  } else {  // ===== epilogue warps =====
    const int quarter = warp & 3;            // TMEM lanes 32*quarter .. +31 are accessible to this warp
    const int half = (warp - 2) >> 2;        // which half of the tile's columns
    int acc = 0;
    uint32_t aph = 0;
    for (int64_t tile = blockIdx.x; tile < sh.total_tiles; tile += gridDim.x) {
      const int g = (int)(tile / tiles_per_group);
      const int rem = (int)(tile % tiles_per_group);
      const int m0 = (rem / sh.tiles_n) * BM, n0 = (rem % sh.tiles_n) * BN;
      mbar_wait(&tfull_bar[acc], aph);
      tc_fence_after();
      const int r = m0 + quarter * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < BN / 64; ++c) {
        const int col = half * (BN / 2) + c * 32;
        float v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + col), v);
        if (r < sh.M && n0 + col < sh.N) epilogue_row32(epi, g, r, n0 + col, v);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if ((acc ^= 1) == 0) aph ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}

// ---- host side ----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  }
  return fn;
}

// 3-D map over a [groups, rows, K] K-contiguous operand; box = 128 bytes of K x box_rows rows.
int make_map(CUtensorMap* map, int dtype, const void* base, int64_t rows, int64_t K, int64_t groups, int64_t gs,
             int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) SVIT_FAIL(SVIT_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const int es = dtype_size(dtype);
  CUtensorMapDataType dt = dtype == SVIT_F32    ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                           : dtype == SVIT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)groups};
  cuuint64_t gstr[2] = {(cuuint64_t)K * es, (cuuint64_t)(groups > 1 ? gs : rows * K) * es};
  cuuint32_t box[3] = {(cuuint32_t)(128 / es), (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, dt, 3, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) SVIT_FAIL(SVIT_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return SVIT_OK;
}

template <int BN, int KIND>
int launch_tc(const CUtensorMap& ma, const CUtensorMap& mb, TcShape sh, const EpiArgs& epi, uint32_t idesc,
              cudaStream_t stream) {
  using C = Cfg<BN>;
  sh.tiles_m = (sh.M + BM - 1) / BM;
  sh.tiles_n = (sh.N + BN - 1) / BN;
  sh.total_tiles = (int64_t)sh.G * sh.tiles_m * sh.tiles_n;
  auto kern = gemm_tc_kernel<BN, KIND>;
  SVIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
  const int64_t grid = std::min<int64_t>(sh.total_tiles, sm_count());
  kern<<<(unsigned)grid, kThreads, C::SMEM, stream>>>(ma, mb, sh, epi, idesc);
  SVIT_LAUNCH_CHECK("gemm_tc_kernel");
  return SVIT_OK;
}

}  // namespace

int gemm_tc(int precision, const void* A, int64_t a_gs, const void* B, int64_t b_gs, int G, int M, int N, int K,
            const EpiArgs& epi, cudaStream_t stream) {
  const int dtype = precision == SVIT_PREC_TF32 ? SVIT_F32 : precision == SVIT_PREC_BF16 ? SVIT_BF16 : SVIT_F16;
  SVIT_CHECK_ARG(precision == SVIT_PREC_TF32 || precision == SVIT_PREC_BF16 || precision == SVIT_PREC_F16,
                 "gemm_tc: precision %d has no tensor-core path", precision);
  const int es = dtype_size(dtype);
  if (!aligned16(A) || !aligned16(B) || ((int64_t)K * es) % 16 || (a_gs * es) % 16 || (b_gs * es) % 16)
    SVIT_FAIL(SVIT_ERR_ALIGN, "gemm_tc: operands must be 16-byte aligned with K*elt and group strides multiples of 16 bytes");
  if (epi.bias && !aligned16(epi.bias)) SVIT_FAIL(SVIT_ERR_ALIGN, "gemm_tc: bias must be 16-byte aligned");
  const int BN = (N % 256 == 0) ? 256 : 128;
  CUtensorMap ma, mb;
  int rc;
  if ((rc = make_map(&ma, dtype, A, M, K, a_gs ? G : 1, a_gs, BM))) return rc;
  if ((rc = make_map(&mb, dtype, B, N, K, b_gs ? G : 1, b_gs, BN))) return rc;
  TcShape sh{};
  sh.G = G, sh.M = M, sh.N = N, sh.K = K;
  const int bk = 128 / es;
  sh.num_kb = (K + bk - 1) / bk;
  sh.a_grouped = a_gs ? 1 : 0;
  SVIT_CHECK_ARG(b_gs != 0 || G == 1, "gemm_tc: B must be grouped when G > 1");
  // instruction descriptor: D fp32, A/B format, both K-major, N, M
  const uint32_t fmt = precision == SVIT_PREC_TF32 ? 2u : precision == SVIT_PREC_BF16 ? 1u : 0u;
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  if (precision == SVIT_PREC_TF32)
    return BN == 256 ? launch_tc<256, 1>(ma, mb, sh, epi, idesc, stream) : launch_tc<128, 1>(ma, mb, sh, epi, idesc, stream);
  return BN == 256 ? launch_tc<256, 0>(ma, mb, sh, epi, idesc, stream) : launch_tc<128, 0>(ma, mb, sh, epi, idesc, stream);
}

}  // namespace svit
