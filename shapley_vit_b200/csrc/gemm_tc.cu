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
// Operand precisions: bf16 / fp16 (kind::f16, K=16 per MMA), tf32 (kind::tf32, K=8), and the two split
// precisions on packed operand arrays (include/svit.h):
//   F16X3  the K loop runs three passes over the fp16 planes: hi*lo, lo*hi, hi*hi (small terms first)
//   F16C8  two e4m3 compensation passes (kind::f8f6f4, K=32 per MMA, twice the f16 rate): hi8*lo8, lo8*hi8,
//          then the fp16 pass, whose FIRST MMA folds the compensation in with scale-input-d:
//          D = A*B + D * 2^-15.  One accumulator, one epilogue, 2 instead of 3 f16-pass equivalents.
//
// Two kernels share the roles above:
//   gemm_tc_kernel   one CTA per 128 x BN tile (cta_group::1); any N (BN = 128 | 256)
//   gemm_tc2_kernel  a CTA PAIR (cluster of 2, tcgen05 cta_group::2) per 256 x 256 tile: each CTA
//                    stages its own 128 rows of A and HALF of the B tile, the leader's MMA thread
//                    issues 256 x 256 x 16 instructions that read both CTAs' shared memory and
//                    write both CTAs' TMEM.  Per CTA and k-block that is 32 KB of L2 -> SM traffic
//                    instead of 48 KB for the same 128 x 256 outputs: with K = 768 the 1-CTA kernel
//                    is bound by L2 -> SM bandwidth (96 B/clk/SM wanted), not by the tensor pipe.
//                    Used whenever N % 256 == 0 and M >= 256.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "epilogue.cuh"
#include "tma_util.h"

namespace svit {
namespace {

constexpr int BM = 128;
constexpr int kEpiWarps = 8;
constexpr int kThreads = (2 + kEpiWarps) * 32;
constexpr int kStagingPerWarp = 32 * 128;  // one 32-row x 32-column fp32 chunk per epilogue warp
constexpr int kBiasPerWarp = 128 * 4;      // the bias slice of the warp's 128 (BN = 256) columns of a tile

template <int BN>
struct Cfg {
  static constexpr int STAGES = BN == 256 ? 3 : 5;  // (1-CTA kernel: small / ragged-N problems only)
  static constexpr int A_BYTES = BM * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int NUM_BARS = 2 * STAGES + 4;
  static constexpr int STAGING_BYTES = kEpiWarps * (kStagingPerWarp + kBiasPerWarp);
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + STAGING_BYTES + NUM_BARS * 8 + 16 + 1024;
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
// (a long suspend-time hint on try_wait was measured: wake-ups got slower, GEMMs lost 4 %; polling stays)
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
// One lane of a converged warp (the lowest): the loops of the producer and MMA warps stay
// warp-uniform, so their counters and addresses live in uniform registers and the issuing lane
// executes ~10 instead of ~20 SASS instructions per tcgen05.mma (the single-thread form made the
// ISSUE loop, not the tensor pipe, the pacer: 71 % tensor-active in profiles/r02_gemm_plain).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
// ---- cluster / CTA-pair helpers ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_addr` (a shared::cta address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // relaxed: the TMEM reads it publishes are ordered by tcgen05.wait::ld + fence::before_thread_sync;
  // a release at cluster scope would first drain every global store of the epilogue (MEMBAR.GPU)
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load issued by either CTA of a pair; the completion bytes are counted on the barrier at
// shared::cluster address `bar_cluster` (the LEADER's full barrier)
__device__ __forceinline__ void tma_load_3d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1,
                                                 int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// arrive (once the MMAs issued so far complete) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
template <int KIND>
__device__ __forceinline__ void tc_mma_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  if constexpr (KIND == 0) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
  }
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
// e4m3 x e4m3 -> fp32 (the compensation passes of SVIT_PREC_F16C8); same instruction descriptor bits as fp16
__device__ __forceinline__ void tc_mma_f8_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tc_mma_f8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// D = A * B + D * 2^-SVIT_C8_SCALE_D (scale-input-d, kind::f16 only): rescales the accumulated compensation
__device__ __forceinline__ void tc_mma_scaled_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%4, %4, %4, %4, %4, %4, %4, %4}, p, %5;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(z), "n"(SVIT_C8_SCALE_D)
      : "memory");
}
__device__ __forceinline__ void tc_mma_scaled(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%4, %4, %4, %4}, p, %5;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(z), "n"(SVIT_C8_SCALE_D)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (row) i
__device__ __forceinline__ void tmem_ld_32x32_issue(uint32_t taddr, float* v) {
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
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// the same descriptor `bytes` further into shared memory (start-address field is in 16-byte units)
__device__ __forceinline__ uint64_t umma_desc_advance(uint64_t desc, uint32_t bytes) { return desc + (bytes >> 4); }

struct TcShape {
  int G, M, N, K;
  int tiles_m, tiles_n, num_kb;
  int64_t total_tiles;
  int a_grouped;  // 0: A shared by all groups
  int b_grouped;  // 0: B shared by all groups (frozen-base LoRA: one weight matrix for every coalition)
  int mode;       // svit_operand_format of the operands: the schedule of the K loop
  int nk_main;    // k-blocks of one fp16 / tf32 pass (128 bytes of K each)
  int nk_aux;     // C8: k-blocks (128 e4m3 values of K) of one compensation pass
  int nk_ext;     // k-blocks (one per pass) appended from the K-EXTENSION operands, 0 = none
};

// the operand planes (main, aux1, aux2) of A and of B; unused entries repeat the main plane.  ea / eb: the planes of the
// K-extension operands (svit_gemm_ext: out = A B^T + A_ext B_ext^T, K_ext = 64, always grouped) -- a low-rank
// per-group correction on top of a shared B rides the same accumulator as extra k-blocks.
struct TcMaps {
  CUtensorMap a[3], b[3];
  CUtensorMap ea[3], eb[3];
};

// Step kb of the K loop: planes of A and B it multiplies, the K coordinate (elements) of its 128-byte block and
// whether the block comes from the extension operands.  Per pass: the main k-blocks, then nk_ext extension blocks.
//   PLAIN  (A0, B0)
//   X3     (A0 hi, B1 lo), (A1 lo, B0 hi), (A0 hi, B0 hi)                           -- small terms first
//   C8     (A1 hi8, B2 lo8), (A2 lo8, B1 hi8) [kind::f8f6f4, 128 values per block], then (A0, B0) [kind::f16]
template <int KIND>
__device__ __forceinline__ void kstep(const TcShape& sh, int kb, int& ia, int& ib, int& kc, bool& ext) {
  constexpr int BKE = KIND == 0 ? 64 : 32;
  const int ne = sh.nk_ext, len_main = sh.nk_main + ne;
  ia = 0, ib = 0, ext = false;
  int pos = kb;
  if (KIND == 0 && sh.mode == SVIT_FMT_X3) {
    const int p = kb / len_main;
    pos = kb - p * len_main;
    if (p == 0) ib = 1;
    else if (p == 1) ia = 1;
  } else if (KIND == 0 && sh.mode == SVIT_FMT_C8) {
    const int len_aux = sh.nk_aux + ne;
    if (kb < 2 * len_aux) {
      const int p = kb >= len_aux ? 1 : 0;
      pos = kb - p * len_aux;
      ia = p ? 2 : 1, ib = p ? 1 : 2;
      ext = pos >= sh.nk_aux;
      kc = ext ? 0 : pos * 128;  // (the extension holds 64 values: the rest of its 128-byte block reads as zeros)
      return;
    }
    pos = kb - 2 * len_aux;
  }
  ext = pos >= sh.nk_main;
  kc = ext ? (pos - sh.nk_main) * BKE : pos * BKE;
}

// ---- packed fp32x2 arithmetic (FFMA2: two IEEE fp32 FMAs per issue slot on sm_100) ------------
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Exact (erf) GELU of two values without erff():  gelu(x) = x * Phi(x),  Phi(-a) = 2^E(-a) for
// a = |x| with E a degree-5 fit of log2(erfc(a / sqrt 2) / 2), so
//   gelu(x) = max(x, 0) - a * 2^E(-a)
// for either sign.  The fit minimises the error of the RESULT (weight a * Phi(-a) * ln 2 on the exponent
// error, iteratively reweighted least squares on [0, 12]); its leading term keeps E -> -inf for large a.
// |error| <= 6.3e-7 absolute in fp32 arithmetic -- the rounding of the result, the same as the degree-7
// unweighted fit it replaces -- against 1.2e-4 for the fp16 rounding of an output of magnitude 0.25.
// 5 + 1 FFMA2, 2 MUFU.EX2, 2 FMNMX per pair -- erff() costs ~5x the issue slots, which made the
// MLP-up epilogue, not the tensor pipe, the bound of that GEMM.
__device__ __forceinline__ void gelu_pair(float& x0, float& x1) {
  const float2 ma = make_float2(-fabsf(x0), -fabsf(x1));
  float2 p = make_float2(4.732936637e-04f, 4.732936637e-04f);
  p = ffma2(p, ma, make_float2(7.084457064e-03f, 7.084457064e-03f));
  p = ffma2(p, ma, make_float2(5.182715352e-02f, 5.182715352e-02f));
  p = ffma2(p, ma, make_float2(-4.599926678e-01f, -4.599926678e-01f));
  p = ffma2(p, ma, make_float2(1.150787756e+00f, 1.150787756e+00f));
  p = ffma2(p, ma, make_float2(-1.000037635e+00f, -1.000037635e+00f));
  const float2 e = make_float2(ex2_approx(p.x), ex2_approx(p.y));
  const float2 r = ffma2(ma, e, make_float2(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
  x0 = r.x, x1 = r.y;
}

__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}

// ---- epilogue of one 32-row x 32-column accumulator chunk ------------------------------------
// tcgen05.ld hands every thread one ROW (32 consecutive columns).  Storing that layout directly
// makes each warp store touch 32 different cache lines (the LSU, not the tensor pipe, then paces
// the GEMM).  So: bias + GELU in the row layout, transpose through a per-warp swizzled shared
// memory tile (conflict-free both ways), then position-embedding / residual adds and the global
// stores in a COALESCED layout (a quarter-warp covers one contiguous 128-byte row segment).

enum { EPI_DIRECT = 0, EPI_RESIDUAL = 1, EPI_GENERIC = 2, EPI_REDUCE = 3 };  // see epilogue_chunk

// the residual fragment of one chunk in the coalesced layout (8 x float4 per lane), read ahead of use
template <int MODE>
__device__ __forceinline__ void residual_prefetch(const EpiArgs& e, int g, int m_slab, int n, int lane, float4 (&res)[8]) {
  const int jj = lane & 7, rsub = lane >> 3;
  const float* base = e.residual + (size_t)g * e.residual_gs + n + jj * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m_slab + i * 4 + rsub;
    res[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < e.M) {
      const int64_t orow = MODE == EPI_GENERIC ? epi_out_row(e, row) : (int64_t)row;
      res[i] = *reinterpret_cast<const float4*>(base + (size_t)orow * e.N);
    }
  }
}

// Epilogue modes: the kernels are compiled once per mode so that the hot loop of an epilogue warp
// holds only the code it runs (a single generic body overflowed the instruction cache: the
// epilogue warps stalled on instruction fetch, profiles/r05_gemm_proj).
//   EPI_DIRECT    + bias (+ GELU) -> cast -> store; no row remap           (QKV, MLP-up)
//   EPI_RESIDUAL  + bias -> + fp32 residual -> fp32 store; no row remap    (attention out-proj, MLP-down)
//   EPI_GENERIC   everything decided at run time                           (patch embedding, odd callers)
//   EPI_REDUCE    + bias -> out += value by a TMA reduce-add (out IS the residual, fp32, in place): the
//                 residual never travels through the SM; pair kernel only       (out-proj, MLP-down in the forward)

//   v        the thread's 32 accumulators (row m_slab + lane, columns n .. n+31)
//   stg      this warp's staging tile (kStagingPerWarp bytes of shared memory)
//   res      this chunk's residual fragment (residual_prefetch), ignored unless the mode adds one
//   bias_s   this chunk's 32 bias values in shared memory (staged per tile), ignored unless e.bias
template <int MODE, int OFMT>
__device__ __forceinline__ void epilogue_chunk(const EpiArgs& e, int g, int m_slab, int n, float* v, uint8_t* stg,
                                               int lane, const float4 (&res)[8], const float* bias_s) {
  const int N = e.N, M = e.M;
  const bool has_rowvec = MODE == EPI_GENERIC && e.rowvec != nullptr;
  const bool has_res = MODE == EPI_RESIDUAL || (MODE == EPI_GENERIC && e.residual != nullptr);
  const bool remap = MODE == EPI_GENERIC && e.rows_in > 0;
  const bool stage16 = MODE == EPI_RESIDUAL ? false : (e.out_dtype != SVIT_F32 && !has_rowvec && !has_res);
  const int jj = lane & 7, rsub = lane >> 3;  // coalesced fp32 layout: 4 rows x 8 float4 per warp access

  if (e.bias) {  // broadcast reads of the staged bias: no global-load latency on the chunk's critical path
    float4 t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = reinterpret_cast<const float4*>(bias_s)[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 a = fadd2(make_float2(v[4 * i], v[4 * i + 1]), make_float2(t[i].x, t[i].y));
      const float2 b = fadd2(make_float2(v[4 * i + 2], v[4 * i + 3]), make_float2(t[i].z, t[i].w));
      v[4 * i] = a.x, v[4 * i + 1] = a.y, v[4 * i + 2] = b.x, v[4 * i + 3] = b.y;
    }
  }
  if (MODE != EPI_RESIDUAL && e.gelu) {
#pragma unroll
    for (int i = 0; i < 32; i += 2) gelu_pair(v[i], v[i + 1]);
  }

  if (stage16) {
    // 16-bit tile: rows of 64 bytes, 16-byte chunk j of row r at chunk (j ^ ((r >> 1) & 3))
    uint32_t w[16];
    uint32_t x[OFMT == SVIT_FMT_PLAIN ? 1 : 16];  // X3: the lo16 pairs; C8: x[0..7] hi8, x[8..15] lo8 (4 values per word)
    if constexpr (OFMT == SVIT_FMT_X3) {
#pragma unroll
      for (int i = 0; i < 16; ++i) split_x3(v[2 * i], v[2 * i + 1], w[i], x[i]);
    } else if constexpr (OFMT == SVIT_FMT_C8) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        uint16_t h0, l0, h1, l1;
        split_c8(v[2 * i], v[2 * i + 1], w[i], h0, l0);
        split_c8(v[2 * i + 2], v[2 * i + 3], w[i + 1], h1, l1);
        x[i >> 1] = (uint32_t)h0 | ((uint32_t)h1 << 16);
        x[8 + (i >> 1)] = (uint32_t)l0 | ((uint32_t)l1 << 16);
      }
    } else if (e.out_dtype == SVIT_BF16) {
#pragma unroll
      for (int i = 0; i < 16; ++i) w[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) w[i] = pack_f16x2_sat(v[2 * i], v[2 * i + 1]);
    }
    if constexpr (OFMT == SVIT_FMT_X3) {  // the lo16 tile: same layout, 2 KB further on
      uint8_t* lb = stg + 2048 + lane * 64;
      const int sl = (lane >> 1) & 3;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(lb + ((j ^ sl) << 4)) = make_uint4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
    } else if constexpr (OFMT == SVIT_FMT_C8) {  // hi8 / lo8 tiles: rows of 32 bytes, chunk j of row r at (j ^ ((r >> 2) & 1))
      uint8_t* hb = stg + 2048 + lane * 32;
      const int sl = (lane >> 2) & 1;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        *reinterpret_cast<uint4*>(hb + ((j ^ sl) << 4)) = make_uint4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
        *reinterpret_cast<uint4*>(hb + 1024 + ((j ^ sl) << 4)) = make_uint4(x[8 + 4 * j], x[9 + 4 * j], x[10 + 4 * j], x[11 + 4 * j]);
      }
    }
    uint8_t* wbase = stg + lane * 64;
    const int sw = (lane >> 1) & 3;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<uint4*>(wbase + ((j ^ sw) << 4)) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
    __syncwarp();
    const int ch = lane & 3;
    uint4 q[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // 8 rows per access, 4 lanes cover one 64-byte row segment
      const int rr = i * 8 + (lane >> 2);
      q[i] = *reinterpret_cast<const uint4*>(stg + rr * 64 + ((ch ^ ((rr >> 1) & 3)) << 4));
    }
    uint16_t* out = reinterpret_cast<uint16_t*>(e.out) + (size_t)g * e.out_gs + n + ch * 8;
    if (remap) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = m_slab + i * 8 + (lane >> 2);
        if (row < M) *reinterpret_cast<uint4*>(out + (size_t)epi_out_row(e, row) * N) = q[i];
      }
    } else {
      uint16_t* o = out + (size_t)(m_slab + (lane >> 2)) * N;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (m_slab + i * 8 + (lane >> 2) < M) *reinterpret_cast<uint4*>(o + (size_t)(i * 8) * N) = q[i];
    }
    if constexpr (OFMT == SVIT_FMT_X3) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = i * 8 + (lane >> 2);
        q[i] = *reinterpret_cast<const uint4*>(stg + 2048 + rr * 64 + ((ch ^ ((rr >> 1) & 3)) << 4));
      }
      uint16_t* lo = reinterpret_cast<uint16_t*>(e.out_operand().aux1()) + (size_t)g * e.out_gs + n + ch * 8;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = m_slab + i * 8 + (lane >> 2);
        if (row < M) *reinterpret_cast<uint4*>(lo + (size_t)(remap ? epi_out_row(e, row) : (int64_t)row) * N) = q[i];
      }
    } else if constexpr (OFMT == SVIT_FMT_C8) {  // 2 lanes cover one 32-byte row segment, 16 rows per access
      const int c8 = lane & 1;
      const Operand oo = e.out_operand();
#pragma unroll
      for (int pl = 0; pl < 2; ++pl) {
        uint4 r8[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int rr = i * 16 + (lane >> 1);
          r8[i] = *reinterpret_cast<const uint4*>(stg + 2048 + pl * 1024 + rr * 32 + ((c8 ^ ((rr >> 2) & 1)) << 4));
        }
        uint8_t* pb = reinterpret_cast<uint8_t*>(pl == 0 ? oo.aux1() : oo.aux2()) + (size_t)g * e.out_gs + n + c8 * 16;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int row = m_slab + i * 16 + (lane >> 1);
          if (row < M) *reinterpret_cast<uint4*>(pb + (size_t)(remap ? epi_out_row(e, row) : (int64_t)row) * N) = r8[i];
        }
      }
    }
  } else {
    // fp32 tile: rows of 128 bytes, 16-byte chunk j of row r at chunk (j ^ (r & 7))
    uint8_t* wbase = stg + lane * 128;
    const int sw = lane & 7;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<float4*>(wbase + ((j ^ sw) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncwarp();
    const size_t base = (size_t)g * e.out_gs + n + jj * 4;
#pragma unroll
    for (int hb = 0; hb < 2; ++hb) {  // two batches of 4 row groups: loads first, then adds + stores
      float4 q[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int rr = (hb * 4 + k) * 4 + rsub;
        q[k] = *reinterpret_cast<const float4*>(stg + rr * 128 + ((jj ^ (rr & 7)) << 4));
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = hb * 4 + k;
        const int row = m_slab + i * 4 + rsub;
        if (row < M) {
          const int64_t orow = remap ? epi_out_row(e, row) : (int64_t)row;
          if (has_rowvec) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(e.rowvec + (size_t)g * e.rowvec_gs +
                                                                   (size_t)(orow % e.rows_out) * N + n + jj * 4));
            q[k].x += t.x, q[k].y += t.y, q[k].z += t.z, q[k].w += t.w;
          }
          if (has_res) q[k].x += res[i].x, q[k].y += res[i].y, q[k].z += res[i].z, q[k].w += res[i].w;
          const size_t idx = base + (size_t)orow * N;
          if (MODE == EPI_RESIDUAL || e.out_dtype == SVIT_F32)
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + idx) = q[k];
          else if (e.out_dtype == SVIT_BF16)
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(e.out) + idx) =
                make_uint2(pack_bf16x2(q[k].x, q[k].y), pack_bf16x2(q[k].z, q[k].w));
          else
            *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(e.out) + idx) =
                make_uint2(pack_f16x2_sat(q[k].x, q[k].y), pack_f16x2_sat(q[k].z, q[k].w));
        }
      }
    }
  }
  __syncwarp();  // the staging tile is rewritten by the next chunk
}

// EPI_REDUCE: out[rows, n .. n+31] += acc + bias, as one TMA reduce-add of the staged fp32 chunk.
// fl(x + fl(acc + bias)) is the same two roundings the reference performs (dense output, then the
// residual add, HF modeling_vit.py:265-268 / :337); the addition happens in the L2, so the residual
// stream is neither loaded into nor stored from the SM, and nothing in the epilogue waits on HBM.
// `stg` alternates between two tiles per warp: the reduce issued two chunks ago must have read its
// tile before it is rewritten (cp.async.bulk.wait_group.read 1).  Rows >= M are clipped by the map.
template <int RBUF>
__device__ __forceinline__ void epilogue_chunk_reduce(const EpiArgs& e, const CUtensorMap* map_out, int g, int m_slab, int n,
                                                      float* v, uint8_t* stg, int lane, const float* bias_s) {
  if (e.bias) {
    float4 t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = reinterpret_cast<const float4*>(bias_s)[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 a = fadd2(make_float2(v[4 * i], v[4 * i + 1]), make_float2(t[i].x, t[i].y));
      const float2 b = fadd2(make_float2(v[4 * i + 2], v[4 * i + 3]), make_float2(t[i].z, t[i].w));
      v[4 * i] = a.x, v[4 * i + 1] = a.y, v[4 * i + 2] = b.x, v[4 * i + 3] = b.y;
    }
  }
  if (lane == 0) {
    if (RBUF == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  __syncwarp();
  // fp32 tile, rows of 128 bytes, 16-byte chunk j of row r at chunk (j ^ (r & 7)) = TMA SWIZZLE_128B
  uint8_t* wbase = stg + lane * 128;
  const int sw = lane & 7;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(wbase + ((j ^ sw) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) {
    asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map_out),
                 "r"(smem_u32(stg)), "r"(n), "r"(m_slab), "r"(g)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
}

// ragged N (not a multiple of 8, or a partial last chunk): element-wise in the row layout
__device__ __forceinline__ void epilogue_row_ragged(const EpiArgs& e, int g, int r, int n, const float* v) {
  const int64_t orow = epi_out_row(e, r);
#pragma unroll
  for (int i = 0; i < 32; ++i)  // static indices keep v[] in registers
    if (n + i < e.N) epi_store(e, g, orow, n + i, epi_apply(e, g, orow, n + i, v[i]));
}

// ---- epilogue of one 128 x BN accumulator tile by one epilogue warp ---------------------------
// The warp owns TMEM lanes 32*quarter.. (rows m_slab..m_slab+31) and columns half*BN/2.. of the tile.
// Everything that comes from global memory is requested BEFORE it is needed: the tile's bias slice
// goes to shared memory and the first residual fragment to registers before the wait on the
// accumulator, each further residual fragment one chunk ahead.  (With the loads issued where they
// are used, an L2 round trip per chunk sat on the epilogue's critical path and the tensor pipe
// idled a third of the time: profiles/r03_gemm_qkv.)
template <int BN, int MODE, int OFMT, int RBUF = 2, class Arrive>
__device__ __forceinline__ void epilogue_tile(const EpiArgs& epi, const TcShape& sh, int g, int m_slab, int n0, int half,
                                              int lane, uint32_t tmem_acc, uint8_t* stg, float* bias_s,
                                              uint64_t* tfull, uint32_t parity, Arrive arrive,
                                              const CUtensorMap* map_out = nullptr) {
  constexpr int NCH = BN / 64;  // 32-column chunks per warp (2 or 4)
  const bool vec_ok = (sh.N & (OFMT == SVIT_FMT_C8 ? 15 : 7)) == 0;  // 16-byte pieces of every output plane
  const int n_w = n0 + half * (BN / 2);
  const bool rows_ok = m_slab < sh.M;
  if (epi.bias && vec_ok) {
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane * 4 < BN / 2 && n_w + lane * 4 < sh.N)
      b = __ldg(reinterpret_cast<const float4*>(epi.bias + (size_t)g * epi.bias_gs + n_w + lane * 4));
    reinterpret_cast<float4*>(bias_s)[lane] = b;
    __syncwarp();
  }
  const bool use_res = (MODE == EPI_RESIDUAL || (MODE == EPI_GENERIC && epi.residual)) && vec_ok && rows_ok;  // (EPI_REDUCE: no)
  float4 res0[8], res1[8];
  if (use_res && n_w + 32 <= sh.N) residual_prefetch<MODE>(epi, g, m_slab, n_w, lane, res0);
  mbar_wait(tfull, parity);
  tc_fence_after();
  // Two chunks per trip, so that the TMEM load and the residual reads of the next chunk are in
  // flight under the math and the stores of the current one.
  float v0[32], v1[32];
  const uint32_t t0 = tmem_acc + (uint32_t)(half * (BN / 2));
  tmem_ld_32x32_issue(t0, v0);
  auto chunk = [&](int n, float* v, const float4 (&res)[8], int c) {
    if (MODE == EPI_REDUCE) {
      if (rows_ok) epilogue_chunk_reduce<RBUF>(epi, map_out, g, m_slab, n, v, stg + (c & (RBUF - 1)) * kStagingPerWarp, lane, bias_s + 32 * c);
      return;
    }
    if (n < sh.N && rows_ok) {
      if (vec_ok && n + 32 <= sh.N)
        epilogue_chunk<MODE, OFMT>(epi, g, m_slab, n, v, stg, lane, res, bias_s + 32 * c);
      else if (m_slab + lane < sh.M)
        epilogue_row_ragged(epi, g, m_slab + lane, n, v);
    }
  };
#pragma unroll 1
  for (int c = 0; c < NCH; c += 2) {
    const int n = n_w + c * 32;
    tmem_ld_wait();
    tmem_ld_32x32_issue(t0 + (uint32_t)((c + 1) * 32), v1);
    if (use_res && n + 64 <= sh.N) residual_prefetch<MODE>(epi, g, m_slab, n + 32, lane, res1);
    chunk(n, v0, res0, c);
    tmem_ld_wait();
    if (c + 2 < NCH) {
      tmem_ld_32x32_issue(t0 + (uint32_t)((c + 2) * 32), v0);
      if (use_res && n + 96 <= sh.N) residual_prefetch<MODE>(epi, g, m_slab, n + 64, lane, res0);
    } else {  // accumulator drained into registers: hand the TMEM buffer back early
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive();
    }
    chunk(n + 32, v1, res1, c + 1);
  }
}

// ---- the kernel --------------------------------------------------------------------------
template <int BN, int KIND, int MODE, int OFMT>
__global__ void __launch_bounds__(kThreads, 1)
    gemm_tc_kernel(const __grid_constant__ TcMaps maps, const TcShape sh, const EpiArgs epi, const uint32_t idesc) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // 1 KB aligned, still a __shared__ pointer
  uint8_t* staging = smem + (size_t)C::STAGES * C::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + C::STAGING_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tfull_bar = empty_bar + C::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a[0]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b[0]) : "memory");
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

  if (warp == 0) {  // ===== TMA producer (warp-uniform loop, one elected lane issues) =====
    int s = 0;
    uint32_t ph = 0;
    for (int64_t tile = blockIdx.x; tile < sh.total_tiles; tile += gridDim.x) {
      const int g = (int)(tile / tiles_per_group);
      const int rem = (int)(tile % tiles_per_group);
      const int m0 = (rem / sh.tiles_n) * BM, n0 = (rem % sh.tiles_n) * BN;
      for (int kb = 0; kb < sh.num_kb; ++kb) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + (size_t)s * C::STAGE_BYTES;
          mbar_expect_tx(&full_bar[s], C::STAGE_BYTES);
          int ia, ib, kc;
          bool ext;
          kstep<KIND>(sh, kb, ia, ib, kc, ext);
          tma_load_3d(sa, ext ? &maps.ea[ia] : &maps.a[ia], &full_bar[s], kc, m0, (ext || sh.a_grouped) ? g : 0);
          tma_load_3d(sa + C::A_BYTES, ext ? &maps.eb[ib] : &maps.b[ib], &full_bar[s], kc, n0, (ext || sh.b_grouped) ? g : 0);
        }
        __syncwarp();
        if (++s == C::STAGES) s = 0, ph ^= 1;
      }
    }
  } else if (warp == 1) {  // ===== MMA issuer (warp-uniform loop, one elected lane issues) =====
    int s = 0, acc = 0;
    uint32_t ph = 0, aph = 0;
    const uint64_t adesc0 = umma_desc(smem_u32(smem)), bdesc0 = umma_desc(smem_u32(smem) + C::A_BYTES);
    const int naux = (KIND == 0 && sh.mode == SVIT_FMT_C8) ? 2 * (sh.nk_aux + sh.nk_ext) : 0;  // e4m3 compensation k-blocks come first
    for (int64_t tile = blockIdx.x; tile < sh.total_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[acc], aph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int kb = 0; kb < sh.num_kb; ++kb) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad = umma_desc_advance(adesc0, (uint32_t)s * C::STAGE_BYTES);
          const uint64_t bd = umma_desc_advance(bdesc0, (uint32_t)s * C::STAGE_BYTES);
          if (KIND == 0 && kb < naux) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_f8(d_tmem, umma_desc_advance(ad, k * 32), umma_desc_advance(bd, k * 32), idesc, (uint32_t)(kb | k));
          } else if (KIND == 0 && naux && kb == naux) {  // first fp16 block: D = A*B + D * 2^-15
            tc_mma_scaled(d_tmem, ad, bd, idesc);
#pragma unroll
            for (int k = 1; k < 4; ++k)
              tc_mma<KIND>(d_tmem, umma_desc_advance(ad, k * 32), umma_desc_advance(bd, k * 32), idesc, 1u);
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)  // 4 x 32 bytes along K inside the 128-byte swizzle span
              tc_mma<KIND>(d_tmem, umma_desc_advance(ad, k * 32), umma_desc_advance(bd, k * 32), idesc, (uint32_t)(kb | k));
          }
          tc_commit(&empty_bar[s]);  // smem stage reusable once these MMAs have read it
          if (kb == sh.num_kb - 1) tc_commit(&tfull_bar[acc]);  // accumulator complete
        }
        __syncwarp();
        if (++s == C::STAGES) s = 0, ph ^= 1;
      }
      if ((acc ^= 1) == 0) aph ^= 1;
    }
  } else {  // ===== epilogue warps =====
    const int quarter = warp & 3;            // TMEM lanes 32*quarter .. +31 are accessible to this warp
    const int half = (warp - 2) >> 2;        // which half of the tile's columns
    uint8_t* stg = staging + (size_t)(warp - 2) * kStagingPerWarp;
    float* bias_s = reinterpret_cast<float*>(staging + (size_t)kEpiWarps * kStagingPerWarp + (size_t)(warp - 2) * kBiasPerWarp);
    int acc = 0;
    uint32_t aph = 0;
    for (int64_t tile = blockIdx.x; tile < sh.total_tiles; tile += gridDim.x) {
      const int g = (int)(tile / tiles_per_group);
      const int rem = (int)(tile % tiles_per_group);
      const int m0 = (rem / sh.tiles_n) * BM, n0 = (rem % sh.tiles_n) * BN;
      uint64_t* te = &tempty_bar[acc];
      epilogue_tile<BN, MODE, OFMT>(epi, sh, g, m0 + quarter * 32, n0, half, lane,
                        tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN), stg, bias_s, &tfull_bar[acc], aph,
                        [te] { mbar_arrive(te); });
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

// ---- the CTA-pair kernel (cta_group::2) ----------------------------------------------------
template <int STAGES_, int STG_PER_WARP = kStagingPerWarp>
struct Cfg2 {
  static constexpr int kStg = STG_PER_WARP;
  static constexpr int BN = 256;                    // N of the pair tile; each CTA stages BN / 2 rows of B
  static constexpr int STAGES = STAGES_;
  static constexpr int A_BYTES = BM * 128;          // this CTA's 128 rows of A, one 128-byte k-block
  static constexpr int B_BYTES = (BN / 2) * 128;    // this CTA's half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int NUM_BARS = 2 * STAGES + 4;
  static constexpr int STAGING_BYTES = kEpiWarps * (STG_PER_WARP + kBiasPerWarp);
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + STAGING_BYTES + NUM_BARS * 8 + 16 + 1024;
};
// the reduce mode double-buffers its staging tile (a TMA reduce reads it asynchronously) and gives up a stage for it
template <int MODE, int STAGES> struct PairCfg { using type = Cfg2<5>; };
template <> struct PairCfg<EPI_REDUCE, 4> { using type = Cfg2<4, 2 * kStagingPerWarp>; };
// ... or keeps all five stages with ONE staging tile per warp (the warp then waits for the previous reduce to have
// read the tile before rewriting it)
template <> struct PairCfg<EPI_REDUCE, 5> { using type = Cfg2<5, kStagingPerWarp>; };

template <int KIND, int STAGES, int MODE, int OFMT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    gemm_tc2_kernel(const __grid_constant__ TcMaps maps, const __grid_constant__ CUtensorMap tma_out, const TcShape sh,
                    const EpiArgs epi, const uint32_t idesc) {
  using C = typename PairCfg<MODE, STAGES>::type;
  static_assert(C::STAGES == STAGES, "stage count is a function of the epilogue mode");
  constexpr int BN = C::BN;
  extern __shared__ uint8_t smem_raw[];
  // identical carve-up in both CTAs: the pair MMA and the multicast commits address the peer's
  // shared memory by the SAME offsets
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // 1 KB aligned, still a __shared__ pointer
  uint8_t* staging = smem + (size_t)C::STAGES * C::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + C::STAGING_BYTES);  // used in the leader only
  uint64_t* empty_bar = full_bar + C::STAGES;                                    // per CTA
  uint64_t* tfull_bar = empty_bar + C::STAGES;                                   // per CTA
  uint64_t* tempty_bar = tfull_bar + 2;                                          // used in the leader only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader (issues the MMAs)

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a[0]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b[0]) : "memory");
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);   // the leader's producer arrives once and expects both CTAs' bytes
      mbar_init(&empty_bar[s], 1);  // one multicast tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);               // one multicast tcgen05.commit
      mbar_init(&tempty_bar[a], 2 * kEpiWarps);  // the epilogue warps of BOTH CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // one warp of each CTA, same warp id, same destination offset
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t tiles_per_group = (int64_t)sh.tiles_m * sh.tiles_n;
  const int64_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp == 0) {  // ===== TMA producer (both CTAs; warp-uniform loop, one elected lane issues) =====
    const uint32_t full0 = mapa_u32(smem_u32(&full_bar[0]), 0);  // the leader's full barriers
    int s = 0;
    uint32_t ph = 0;
    for (int64_t tile = pair; tile < sh.total_tiles; tile += npairs) {
      const int g = (int)(tile / tiles_per_group);
      const int rem = (int)(tile % tiles_per_group);
      const int m0 = (rem / sh.tiles_n) * (2 * BM) + (int)rank * BM;
      const int n0 = (rem % sh.tiles_n) * BN + (int)rank * (BN / 2);
      for (int kb = 0; kb < sh.num_kb; ++kb) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + (size_t)s * C::STAGE_BYTES;
          if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * C::STAGE_BYTES);
          int ia, ib, kc;
          bool ext;
          kstep<KIND>(sh, kb, ia, ib, kc, ext);
          tma_load_3d_pair(sa, ext ? &maps.ea[ia] : &maps.a[ia], full0 + 8u * s, kc, m0, (ext || sh.a_grouped) ? g : 0);
          tma_load_3d_pair(sa + C::A_BYTES, ext ? &maps.eb[ib] : &maps.b[ib], full0 + 8u * s, kc, n0,
                           (ext || sh.b_grouped) ? g : 0);
        }
        __syncwarp();
        if (++s == C::STAGES) s = 0, ph ^= 1;
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {  // ===== MMA issuer (leader CTA only; warp-uniform loop, one elected lane issues) =====
      int s = 0, acc = 0;
      uint32_t ph = 0, aph = 0;
      const uint64_t adesc0 = umma_desc(smem_u32(smem)), bdesc0 = umma_desc(smem_u32(smem) + C::A_BYTES);
      const int naux = (KIND == 0 && sh.mode == SVIT_FMT_C8) ? 2 * (sh.nk_aux + sh.nk_ext) : 0;  // e4m3 compensation k-blocks come first
      for (int64_t tile = pair; tile < sh.total_tiles; tile += npairs) {
        mbar_wait(&tempty_bar[acc], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < sh.num_kb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t ad = umma_desc_advance(adesc0, (uint32_t)s * C::STAGE_BYTES);
            const uint64_t bd = umma_desc_advance(bdesc0, (uint32_t)s * C::STAGE_BYTES);
            if (KIND == 0 && kb < naux) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                tc_mma_f8_pair(d_tmem, umma_desc_advance(ad, k * 32), umma_desc_advance(bd, k * 32), idesc, (uint32_t)(kb | k));
            } else if (KIND == 0 && naux && kb == naux) {  // first fp16 block: D = A*B + D * 2^-15
              tc_mma_scaled_pair(d_tmem, ad, bd, idesc);
#pragma unroll
              for (int k = 1; k < 4; ++k)
                tc_mma_pair<KIND>(d_tmem, umma_desc_advance(ad, k * 32), umma_desc_advance(bd, k * 32), idesc, 1u);
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                tc_mma_pair<KIND>(d_tmem, umma_desc_advance(ad, k * 32), umma_desc_advance(bd, k * 32), idesc, (uint32_t)(kb | k));
            }
            tc_commit_pair(&empty_bar[s]);  // frees stage s in both CTAs
            if (kb == sh.num_kb - 1) tc_commit_pair(&tfull_bar[acc]);  // both CTAs' halves of the accumulator are complete
          }
          __syncwarp();
          if (++s == C::STAGES) s = 0, ph ^= 1;
        }
        if ((acc ^= 1) == 0) aph ^= 1;
      }
    }
  } else {  // ===== epilogue warps (both CTAs, each its own 128 rows) =====
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    uint8_t* stg = staging + (size_t)(warp - 2) * C::kStg;
    float* bias_s = reinterpret_cast<float*>(staging + (size_t)kEpiWarps * C::kStg + (size_t)(warp - 2) * kBiasPerWarp);
    const uint32_t tempty0 = mapa_u32(smem_u32(&tempty_bar[0]), 0);
    int acc = 0;
    uint32_t aph = 0;
    for (int64_t tile = pair; tile < sh.total_tiles; tile += npairs) {
      const int g = (int)(tile / tiles_per_group);
      const int rem = (int)(tile % tiles_per_group);
      const int m0 = (rem / sh.tiles_n) * (2 * BM) + (int)rank * BM, n0 = (rem % sh.tiles_n) * BN;
      const uint32_t te = tempty0 + 8u * acc;
      epilogue_tile<BN, MODE, OFMT, C::kStg / kStagingPerWarp>(epi, sh, g, m0 + quarter * 32, n0, half, lane,
                        tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN), stg, bias_s, &tfull_bar[acc], aph,
                        [te] { mbar_arrive_cluster(te); }, &tma_out);
      if ((acc ^= 1) == 0) aph ^= 1;
    }
    if (MODE == EPI_REDUCE && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all reduce-adds performed
  }
  // Nobody leaves while the peer may still multicast a commit into, or read operands from, this
  // CTA's shared memory; TMEM is released by both CTAs after that.
  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}

// ---- host side ----------------------------------------------------------------------------
// 3-D map over a [groups, rows, K] K-contiguous operand plane; box = 128 bytes of K x box_rows rows.
int make_map(CUtensorMap* map, int dtype, const void* base, int64_t rows, int64_t K, int64_t groups, int64_t gs,
             int box_rows) {
  const int es = dtype_size(dtype);
  return encode_map_3d(map, dtype, base, (uint64_t)K, (uint64_t)rows, (uint64_t)groups, (uint64_t)K * es,
                       (uint64_t)(groups > 1 ? gs : rows * K) * es, (uint32_t)(128 / es), (uint32_t)box_rows);
}

// the maps of every plane of one operand (unused entries repeat the main plane)
int make_plane_maps(CUtensorMap (&maps)[3], int dtype, const Operand& op, int64_t off, int64_t rows, int64_t K, int64_t groups,
                    int64_t gs, int box_rows) {
  int rc;
  const int es = dtype_size(dtype);
  if ((rc = make_map(&maps[0], dtype, op.plane(0, off, es), rows, K, groups, gs, box_rows))) return rc;
  maps[1] = maps[0], maps[2] = maps[0];
  if (op.fmt == SVIT_FMT_X3) {
    if ((rc = make_map(&maps[1], SVIT_F16, op.plane(1, off), rows, K, groups, gs, box_rows))) return rc;
  } else if (op.fmt == SVIT_FMT_C8) {
    if ((rc = make_map(&maps[1], SVIT_U8, op.plane(1, off), rows, K, groups, gs, box_rows))) return rc;
    if ((rc = make_map(&maps[2], SVIT_U8, op.plane(2, off), rows, K, groups, gs, box_rows))) return rc;
  }
  return SVIT_OK;
}

template <int BN, int KIND, int OFMT>
int launch_tc(const TcMaps& maps, TcShape sh, const EpiArgs& epi, uint32_t idesc, cudaStream_t stream) {
  using C = Cfg<BN>;
  sh.tiles_m = (sh.M + BM - 1) / BM;
  sh.tiles_n = (sh.N + BN - 1) / BN;
  sh.total_tiles = (int64_t)sh.G * sh.tiles_m * sh.tiles_n;
  auto kern = gemm_tc_kernel<BN, KIND, EPI_GENERIC, OFMT>;
  SVIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
  const int64_t grid = std::min<int64_t>(sh.total_tiles, sm_count());
  kern<<<(unsigned)grid, kThreads, C::SMEM, stream>>>(maps, sh, epi, idesc);
  SVIT_LAUNCH_CHECK("gemm_tc_kernel");
  return SVIT_OK;
}

template <int KIND, int MODE, int OFMT, int STAGES = (MODE == EPI_REDUCE ? 4 : 5)>
int launch_tc2s(const TcMaps& maps, TcShape sh, const EpiArgs& epi, uint32_t idesc, cudaStream_t stream) {
  using C = typename PairCfg<MODE, STAGES>::type;
  CUtensorMap mo = maps.a[0];  // only the reduce mode reads it
  if (MODE == EPI_REDUCE) {
    int rc = encode_map_3d(&mo, SVIT_F32, epi.out, (uint64_t)sh.N, (uint64_t)sh.M, (uint64_t)sh.G, (uint64_t)sh.N * 4,
                           (uint64_t)(sh.G > 1 ? epi.out_gs : (int64_t)sh.M * sh.N) * 4, 32, 32);
    if (rc) return rc;
  }
  sh.tiles_m = (sh.M + 2 * BM - 1) / (2 * BM);
  sh.tiles_n = sh.N / C::BN;
  sh.total_tiles = (int64_t)sh.G * sh.tiles_m * sh.tiles_n;
  auto kern = gemm_tc2_kernel<KIND, STAGES, MODE, OFMT>;
  SVIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
  const int64_t npairs = std::min<int64_t>(sh.total_tiles, sm_count() / 2);
  kern<<<(unsigned)(2 * npairs), kThreads, C::SMEM, stream>>>(maps, mo, sh, epi, idesc);
  SVIT_LAUNCH_CHECK("gemm_tc2_kernel");
  return SVIT_OK;
}

// the output format is a compile-time parameter of the epilogue (only the 16-bit store paths differ)
template <int KIND, int MODE>
int launch_tc2f(const TcMaps& maps, const TcShape& sh, const EpiArgs& epi, uint32_t idesc, cudaStream_t stream) {
  if (KIND == 0 && epi.out_fmt == SVIT_FMT_X3) return launch_tc2s<KIND, MODE, KIND == 0 ? SVIT_FMT_X3 : 0>(maps, sh, epi, idesc, stream);
  if (KIND == 0 && epi.out_fmt == SVIT_FMT_C8) return launch_tc2s<KIND, MODE, KIND == 0 ? SVIT_FMT_C8 : 0>(maps, sh, epi, idesc, stream);
  return launch_tc2s<KIND, MODE, SVIT_FMT_PLAIN>(maps, sh, epi, idesc, stream);
}

template <int KIND>
int launch_tc2(const TcMaps& maps, const TcShape& sh, const EpiArgs& epi, uint32_t idesc, cudaStream_t stream) {
  static const bool no_reduce = [] { const char* e = getenv("SVIT_GEMM_NO_REDUCE"); return e && e[0] == '1'; }();
  const bool remap = epi.rows_in > 0;
  if (!epi.rowvec && !epi.residual && !remap) return launch_tc2f<KIND, EPI_DIRECT>(maps, sh, epi, idesc, stream);
  if (!epi.rowvec && epi.residual && !remap && !epi.gelu && epi.out_dtype == SVIT_F32) {
    // in place (out IS the residual): the add can be done by the L2 (TMA reduce-add)
    if (!no_reduce && epi.residual == epi.out && (sh.G == 1 || epi.residual_gs == epi.out_gs)) {
      // default: five stages, one staging tile (A/B in one session, scripts/r2_gpu4.sh: out-proj 462 -> 441 us, MLP-down
      // and the bench step within the box's noise); SVIT_GEMM_REDUCE_STAGES=4 keeps the double-staged variant
      static const bool five = [] { const char* e = getenv("SVIT_GEMM_REDUCE_STAGES"); return !(e && e[0] == '4'); }();
      return five ? launch_tc2s<KIND, EPI_REDUCE, SVIT_FMT_PLAIN, 5>(maps, sh, epi, idesc, stream)
                  : launch_tc2s<KIND, EPI_REDUCE, SVIT_FMT_PLAIN, 4>(maps, sh, epi, idesc, stream);
    }
    return launch_tc2s<KIND, EPI_RESIDUAL, SVIT_FMT_PLAIN>(maps, sh, epi, idesc, stream);
  }
  return launch_tc2f<KIND, EPI_GENERIC>(maps, sh, epi, idesc, stream);
}

template <int BN, int KIND>
int launch_tc1(const TcMaps& maps, const TcShape& sh, const EpiArgs& epi, uint32_t idesc, cudaStream_t stream) {
  if (KIND == 0 && epi.out_fmt == SVIT_FMT_X3) return launch_tc<BN, KIND, KIND == 0 ? SVIT_FMT_X3 : 0>(maps, sh, epi, idesc, stream);
  if (KIND == 0 && epi.out_fmt == SVIT_FMT_C8) return launch_tc<BN, KIND, KIND == 0 ? SVIT_FMT_C8 : 0>(maps, sh, epi, idesc, stream);
  return launch_tc<BN, KIND, SVIT_FMT_PLAIN>(maps, sh, epi, idesc, stream);
}

}  // namespace

int gemm_tc(int precision, const Operand& A, int64_t a_off, int64_t a_gs, const Operand& B, int64_t b_off, int64_t b_gs, int G,
            int M, int N, int K, const EpiArgs& epi, cudaStream_t stream, const GemmExt* ext) {
  const int mode = format_of_precision(precision);  // operand format = schedule of the K loop
  const int dtype = precision == SVIT_PREC_TF32 ? SVIT_F32 : precision == SVIT_PREC_BF16 ? SVIT_BF16 : SVIT_F16;
  SVIT_CHECK_ARG(precision == SVIT_PREC_TF32 || precision == SVIT_PREC_BF16 || precision == SVIT_PREC_F16 || mode != SVIT_FMT_PLAIN,
                 "gemm_tc: precision %d has no tensor-core path", precision);
  SVIT_CHECK_ARG(A.fmt == mode && B.fmt == mode, "gemm_tc: operand formats (%d, %d) do not match precision %d", A.fmt, B.fmt,
                 precision);
  SVIT_CHECK_ARG(epi.out_fmt == SVIT_FMT_PLAIN || (epi.out_dtype == SVIT_F16 && !epi.rowvec && !epi.residual),
                 "gemm_tc: a split-format output takes neither rowvec nor residual");
  const int es = dtype_size(dtype);
  const int kal = mode == SVIT_FMT_C8 ? 16 : 16 / es;  // every plane's rows and group strides must be 16-byte multiples
  if (!aligned16(A.base) || !aligned16(B.base) || K % kal || a_off % 16 || b_off % 16 || a_gs % kal || b_gs % kal ||
      (mode != SVIT_FMT_PLAIN && (A.alloc % 16 || B.alloc % 16)))
    SVIT_FAIL(SVIT_ERR_ALIGN, "gemm_tc: operands must be 16-byte aligned with K, offsets and group strides multiples of 16 bytes in every plane");
  if (N % 8 == 0) {  // the vectorised epilogue moves 16-byte pieces of bias / rowvec / residual / out
    const int oes = dtype_size(epi.out_dtype);
    if ((epi.bias && (!aligned16(epi.bias) || epi.bias_gs % 4)) || (epi.rowvec && (!aligned16(epi.rowvec) || epi.rowvec_gs % 4)) ||
        (epi.residual && (!aligned16(epi.residual) || epi.residual_gs % 4)) || !aligned16(epi.out) || (epi.out_gs * oes) % 16 ||
        (epi.out_fmt != SVIT_FMT_PLAIN && (epi.out_alloc % 16 || epi.out_gs % 16)))
      SVIT_FAIL(SVIT_ERR_ALIGN, "gemm_tc: bias/rowvec/residual/out must be 16-byte aligned with 16-byte group strides");
  }
  const int BN = (N % 256 == 0) ? 256 : 128;
  static const bool no_pair = [] { const char* e = getenv("SVIT_GEMM_NO_PAIR"); return e && e[0] == '1'; }();
  const bool pair = BN == 256 && M >= 2 * BM && !no_pair;  // CTA-pair kernel: each CTA stages half of the B tile
  TcMaps maps;
  int rc;
  if ((rc = make_plane_maps(maps.a, dtype, A, a_off, M, K, a_gs ? G : 1, a_gs, BM))) return rc;
  if ((rc = make_plane_maps(maps.b, dtype, B, b_off, N, K, b_gs ? G : 1, b_gs, pair ? BN / 2 : BN))) return rc;
  for (int i = 0; i < 3; ++i) maps.ea[i] = maps.a[i], maps.eb[i] = maps.b[i];
  if (ext) {
    SVIT_CHECK_ARG(ext->A.fmt == mode && ext->B.fmt == mode && ext->a_gs % 16 == 0 && ext->b_gs % 16 == 0 && ext->a_off % 16 == 0 &&
                       ext->b_off % 16 == 0 && (G == 1 || (ext->a_gs && ext->b_gs)),
                   "gemm_tc: the K-extension operands must be grouped arrays in the precision's format, 16-element aligned");
    if ((rc = make_plane_maps(maps.ea, dtype, ext->A, ext->a_off, M, kGemmExtK, G, ext->a_gs, BM))) return rc;
    if ((rc = make_plane_maps(maps.eb, dtype, ext->B, ext->b_off, N, kGemmExtK, G, ext->b_gs, pair ? BN / 2 : BN))) return rc;
  }
  TcShape sh{};
  sh.G = G, sh.M = M, sh.N = N, sh.K = K;
  const int bk = 128 / es;
  sh.mode = mode;
  sh.nk_main = (K + bk - 1) / bk;
  sh.nk_aux = mode == SVIT_FMT_C8 ? (K + 127) / 128 : 0;
  sh.nk_ext = ext ? kGemmExtK / bk : 0;  // (one k-block per pass; two for tf32)
  sh.num_kb = mode == SVIT_FMT_X3 ? 3 * (sh.nk_main + sh.nk_ext)
                                  : (sh.nk_main + sh.nk_ext) + (mode == SVIT_FMT_C8 ? 2 * (sh.nk_aux + sh.nk_ext) : 0);
  sh.a_grouped = a_gs ? 1 : 0;
  sh.b_grouped = b_gs ? 1 : 0;
  // instruction descriptor: D fp32, A/B format, both K-major, N, M (format 0 is fp16 for kind::f16 and e4m3 for
  // kind::f8f6f4: the compensation passes of F16C8 use the same descriptor)
  const uint32_t fmt = precision == SVIT_PREC_TF32 ? 2u : precision == SVIT_PREC_BF16 ? 1u : 0u;
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) |
                         ((uint32_t)((pair ? 2 * BM : BM) >> 4) << 24);
  if (pair)
    return precision == SVIT_PREC_TF32 ? launch_tc2<1>(maps, sh, epi, idesc, stream) : launch_tc2<0>(maps, sh, epi, idesc, stream);
  if (precision == SVIT_PREC_TF32)
    return BN == 256 ? launch_tc1<256, 1>(maps, sh, epi, idesc, stream) : launch_tc1<128, 1>(maps, sh, epi, idesc, stream);
  return BN == 256 ? launch_tc1<256, 0>(maps, sh, epi, idesc, stream) : launch_tc1<128, 0>(maps, sh, epi, idesc, stream);
}

}  // namespace svit
