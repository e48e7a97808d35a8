// Split-precision fused-softmax attention on the 5th-generation tensor cores (tcgen05 + TMEM):
// the attention of SVIT_PREC_F16X3 / SVIT_PREC_F16C8 for head_dim 64, 128 < T <= 224 (ViT-B/L at 224 px).
//
// Same contract as attention_tc.cu (HF ViTSelfAttention, modeling_vit.py:171-196, 220-252, as the
// reference evaluates it at federated_learning/utils.py:886), but every product carries ~21 bits:
//     S = Ql Kh^T + Qh Kl^T + Qh Kh^T        (12 tcgen05.mma per 128-row query tile, small terms first)
//     O = Pl Vh   + Ph Vl   + Ph Vh          (3 per 16 keys; P = hi + lo written back to TMEM by the softmax warps)
// Q, K, V arrive as the fp16 hi / lo PLANES the QKV GEMM's epilogue emits (SVIT_FMT_X3), by TMA; the context
// leaves through TMA stores in the operand format of the out-projection GEMM (X3 or C8 planes).
//
// Budget of one SM (one persistent CTA): shared memory holds ONE item -- Q (128 + R1 rows, hi + lo), K and V
// (NK rows each, hi + lo: 104 KB at T = 197) and the output staging tiles: 210 KB -- so the smem ring of
// attention_tc.cu is replaced by per-operand full / free barriers: Q and K of item i+1 are requested as soon as
// the score MMAs of item i have read them, V as soon as its P V MMAs have, all hidden behind the softmax.
// TMEM: S0 at column 0, S1 at 224, one 64-column O buffer at 448 shared by the two query tiles (the issue order
// S0(i), O1(i-1), S1(i), O0(i) never has both O live).  P overwrites S chunk by chunk: the hi half of a
// 32-column chunk in its first 16 columns, the lo half in the last 16, so no unread score is ever overwritten.
#include <cuda.h>

#include <algorithm>

#include "elementwise.h"
#include "tma_util.h"

namespace svit {
namespace {

constexpr int kD = 64;
constexpr int kThreads = 10 * 32;
constexpr int kProducerWarp = 8, kMmaWarp = 9;  // (high warp ids: the issue arbiter prefers them)
constexpr uint32_t kS1Col = 224, kOCol = 448;
constexpr size_t kSmemLimit = 232448;
enum { B_QFULL0 = 0, B_QFULL1, B_QFREE0, B_QFREE1, B_KFULL, B_KFREE, B_VFULL, B_VFREE, B_SFULL0, B_SFULL1, B_PFULL0, B_PFULL1,
       B_OFULL0, B_OFULL1, B_SFREE0, B_SFREE1, B_COUNT };

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
// bounded: a broken pipeline traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity))
    if (clock64() - t0 > 4000000000LL) __trap();
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// 128B-swizzled operand tile, rows of 128 bytes, 8-row groups 1024 bytes apart (K-major: rows = M/N index,
// MN-major: rows = K index; the instruction descriptor says which)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

#define TMEM_LD_X32(taddr, r)                                                                                      \
  asm volatile(                                                                                                    \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                    \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                    \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                    \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),      \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),     \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                   \
      : "r"(taddr)                                                                                                 \
      : "memory")
#define TMEM_LD_X16(taddr, r)                                                                                      \
  asm volatile(                                                                                                    \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                                    \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"                             \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                    \
      : "r"(taddr)                                                                                                 \
      : "memory")
#define TMEM_ST_X8(taddr, r)                                                                                       \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),       \
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])               \
               : "memory")
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}

// Debug aid (scripts/attn_split_trace.py; build with -DSVIT_ATT_TRACE): CTA 0 records clock64() per phase of its first
// items.  Compiled out by default: the instrumented kernel is 20 % slower (665 vs 553 us per 1 024 x 12 heads).
__device__ long long* g_att_split_trace = nullptr;
#ifdef SVIT_ATT_TRACE
constexpr int kTraceItems = 24, kTraceSlots = 16;
#define ATT_TRACE(slot)                                                                      \
  do {                                                                                       \
    if (trace && lane == 0 && it < kTraceItems) trace[it * kTraceSlots + (slot)] = clock64(); \
  } while (0)
#else
#define ATT_TRACE(slot) do { } while (0)
#endif

struct AttMaps {
  CUtensorMap q0[2], q1[2], kv[2];  // [hi, lo] planes of qkv: 128-row / R1-row query boxes, NK-row key / value boxes
  CUtensorMap o[3];                 // context planes: hi16, then lo16 (X3) or hi8, lo8 (C8)
};

struct AttShape {
  int T, R1, heads;      // tokens, rows of the second query tile (T - 128 rounded up to 8), heads
  int64_t items;         // n_seq * heads
  float sl2;             // d^-0.5 * log2(e)
  uint32_t idesc_qk, idesc_pv;
  // byte offsets from the 1 KB-aligned start of dynamic shared memory
  uint32_t off_q[2][2];  // [tile][hi, lo]
  uint32_t off_k[2], off_v[2];
  uint32_t off_o[2];     // output staging of the two softmax groups (256 bytes per row, 8 KB per warp)
  uint32_t off_bar;
};

// issue the TMEM load of the 32- (W16: 16-) column chunk at column c0 of the thread's row
template <bool W16>
__device__ __forceinline__ void chunk_load(uint32_t tbuf, int c0, uint32_t (&r)[32]) {
  if (W16) {
    TMEM_LD_X16(tbuf + (uint32_t)c0, r);
  } else {
    TMEM_LD_X32(tbuf + (uint32_t)c0, r);
  }
}

// second pass on one loaded chunk: p = 2^(s * sl2 + noff) accumulated into `sum`, split p = hi + lo (fp16 pairs)
// and written back over the chunk's own columns: hi half first, lo half behind it.  MASK: keys >= Tn give 0.
template <bool W16, bool MASK>
__device__ __forceinline__ void chunk_exp(uint32_t tbuf, int c0, int Tn, float sl2, float noff, const uint32_t (&r)[32],
                                          float2 (&sum)[2]) {
  constexpr int W = W16 ? 16 : 32;
  uint32_t hi[16], lo[16];
  const float2 sc = make_float2(sl2, sl2), of = make_float2(noff, noff);
#pragma unroll
  for (int i = 0; i < W; i += 2) {
    const float2 x = ffma2(make_float2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), sc, of);
    float2 e = make_float2(ex2_approx(x.x), ex2_approx(x.y));
    if (MASK) {
      if (c0 + i >= Tn) e.x = 0.f;
      if (c0 + i + 1 >= Tn) e.y = 0.f;
    }
    sum[(i >> 1) & 1] = fadd2(sum[(i >> 1) & 1], e);
    split_x3(e.x, e.y, hi[i >> 1], lo[i >> 1]);
  }
  TMEM_ST_X8(tbuf + (uint32_t)c0, hi);
  if (!W16) {
    TMEM_ST_X8(tbuf + (uint32_t)c0 + 8, (hi + 8));
    TMEM_ST_X8(tbuf + (uint32_t)c0 + 16, lo);
    TMEM_ST_X8(tbuf + (uint32_t)c0 + 24, (lo + 8));
  } else {
    TMEM_ST_X8(tbuf + (uint32_t)c0 + 8, lo);
  }
}

// first pass on one loaded chunk: running row maxima (four independent chains)
template <bool W16, bool MASK>
__device__ __forceinline__ void chunk_max(int c0, int Tn, const uint32_t (&r)[32], float (&m)[4]) {
  constexpr int W = W16 ? 16 : 32;
#pragma unroll
  for (int i = 0; i < W; i += 2) {
    float a = __uint_as_float(r[i]), b = __uint_as_float(r[i + 1]);
    if (MASK) {
      if (c0 + i >= Tn) a = -INFINITY;
      if (c0 + i + 1 >= Tn) b = -INFINITY;
    }
    m[(i >> 1) & 3] = fmax3(m[(i >> 1) & 3], a, b);
  }
}

// Softmax of the thread's row of S (NK columns at tbuf), P = hi + lo written in place; returns 1 / row sum.
// Both passes run their TMEM loads one chunk ahead of the arithmetic (two register buffers).  Keys >= T can
// only sit in the last two chunks (NK - 48 < T).
template <int NK>
__device__ __forceinline__ float softmax_row(uint32_t tbuf, int Tn, float sl2) {
  constexpr int NFULL = NK / 32;           // 32-column chunks
  constexpr bool TAIL = (NK % 32) != 0;    // plus one 16-column chunk
  constexpr int NCH = NFULL + (TAIL ? 1 : 0);
  uint32_t ra[32], rb[32];
  float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  chunk_load<false>(tbuf, 0, ra);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    uint32_t(&cur)[32] = (c & 1) ? rb : ra;
    uint32_t(&nxt)[32] = (c & 1) ? ra : rb;
    tmem_ld_wait();
    if (c + 1 < NCH) {
      if (TAIL && c + 1 == NCH - 1)
        chunk_load<true>(tbuf, (c + 1) * 32, nxt);
      else
        chunk_load<false>(tbuf, (c + 1) * 32, nxt);
    } else {
      chunk_load<false>(tbuf, 0, nxt);  // pass 2, chunk 0
    }
    if (c == NCH - 1) {
      if (TAIL)
        chunk_max<true, true>(c * 32, Tn, cur, mx);
      else
        chunk_max<false, true>(c * 32, Tn, cur, mx);
    } else if (c == NCH - 2) {
      chunk_max<false, true>(c * 32, Tn, cur, mx);
    } else {
      chunk_max<false, false>(c * 32, Tn, cur, mx);
    }
  }
  const float m = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
  const float noff = -m * sl2;
  float2 sum[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    // pass 1 left its extra load in the buffer pass-1 chunk NCH would have used
    uint32_t(&cur)[32] = ((c + NCH) & 1) ? rb : ra;
    uint32_t(&nxt)[32] = ((c + NCH) & 1) ? ra : rb;
    tmem_ld_wait();
    if (c + 1 < NCH) {
      if (TAIL && c + 1 == NCH - 1)
        chunk_load<true>(tbuf, (c + 1) * 32, nxt);
      else
        chunk_load<false>(tbuf, (c + 1) * 32, nxt);
    }
    if (c == NCH - 1) {
      if (TAIL)
        chunk_exp<true, true>(tbuf, c * 32, Tn, sl2, noff, cur, sum);
      else
        chunk_exp<false, true>(tbuf, c * 32, Tn, sl2, noff, cur, sum);
    } else if (c == NCH - 2) {
      chunk_exp<false, true>(tbuf, c * 32, Tn, sl2, noff, cur, sum);
    } else {
      chunk_exp<false, false>(tbuf, c * 32, Tn, sl2, noff, cur, sum);
    }
  }
  tmem_st_wait();
  return 1.0f / ((sum[0].x + sum[0].y) + (sum[1].x + sum[1].y));
}

// NK: keys padded to the kernel's bucket (compile time: the softmax loops are fully unrolled)
template <int NK, int OFMT>
__global__ void __launch_bounds__(kThreads, 1) attention_tc_split_kernel(const __grid_constant__ AttMaps maps, const AttShape sh) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sh.off_bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef SVIT_ATT_TRACE
  long long* const trace = blockIdx.x == 0 ? g_att_split_trace : nullptr;
#endif

  if (warp == kProducerWarp && lane == 0) {
    for (int p = 0; p < 2; ++p) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.q0[p]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.q1[p]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.kv[p]) : "memory");
    }
    for (int i = 0; i < B_COUNT; ++i) mbar_init(&bars[i], 1);
    mbar_init(&bars[B_PFULL0], 4), mbar_init(&bars[B_PFULL1], 4);
    mbar_init(&bars[B_SFREE0], 4), mbar_init(&bars[B_SFREE1], 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kProducerWarp) {  // ===== TMA producer: Q0, K, Q1, V of every item, each as soon as its buffer is free =====
    int it = 0;
    for (int64_t item = blockIdx.x; item < sh.items; item += gridDim.x, ++it) {
      const uint32_t fpar = ((uint32_t)it & 1) ^ 1;  // the buffer's previous use (item it - 1) has been read
      const int seq = (int)(item / sh.heads), head = (int)(item % sh.heads);
      const int h = sh.heads * kD;
      const int cq = head * kD, ck = h + head * kD, cv = 2 * h + head * kD;
      mbar_wait(&bars[B_QFREE0], fpar);
      if (elect_one()) {
        mbar_expect_tx(&bars[B_QFULL0], 2u * 128 * 128);
        tma_load_3d(smem + sh.off_q[0][0], &maps.q0[0], &bars[B_QFULL0], cq, 0, seq);
        tma_load_3d(smem + sh.off_q[0][1], &maps.q0[1], &bars[B_QFULL0], cq, 0, seq);
      }
      __syncwarp();
      mbar_wait(&bars[B_KFREE], fpar);
      if (elect_one()) {
        mbar_expect_tx(&bars[B_KFULL], 2u * NK * 128);
        tma_load_3d(smem + sh.off_k[0], &maps.kv[0], &bars[B_KFULL], ck, 0, seq);
        tma_load_3d(smem + sh.off_k[1], &maps.kv[1], &bars[B_KFULL], ck, 0, seq);
      }
      __syncwarp();
      mbar_wait(&bars[B_QFREE1], fpar);
      if (elect_one()) {
        mbar_expect_tx(&bars[B_QFULL1], 2u * (uint32_t)sh.R1 * 128);
        tma_load_3d(smem + sh.off_q[1][0], &maps.q1[0], &bars[B_QFULL1], cq, 128, seq);
        tma_load_3d(smem + sh.off_q[1][1], &maps.q1[1], &bars[B_QFULL1], cq, 128, seq);
      }
      __syncwarp();
      mbar_wait(&bars[B_VFREE], fpar);
      if (elect_one()) {
        mbar_expect_tx(&bars[B_VFULL], 2u * NK * 128);
        tma_load_3d(smem + sh.off_v[0], &maps.kv[0], &bars[B_VFULL], cv, 0, seq);
        tma_load_3d(smem + sh.off_v[1], &maps.kv[1], &bars[B_VFULL], cv, 0, seq);
      }
      __syncwarp();
    }
  } else if (warp == kMmaWarp) {  // ===== MMA issuer =====
    // Issue order per item i:  S0(i), O1(i-1), S1(i), O0(i): the softmax groups run out of step, and the single
    // O buffer is never claimed twice: S_b(i) waits for group b to have drained O_b(i-1).
    constexpr uint32_t nks = (uint32_t)NK / 16;
    const uint32_t sb = smem_u32(smem);
    auto issue_qk = [&](int b, uint32_t par) {  // S_b = Ql Kh^T + Qh Kl^T + Qh Kh^T
      mbar_wait(&bars[B_SFREE0 + b], par ^ 1);
      mbar_wait(&bars[B_QFULL0 + b], par);
      if (b == 0) mbar_wait(&bars[B_KFULL], par);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t qh = umma_desc(sb + sh.off_q[b][0]), ql = umma_desc(sb + sh.off_q[b][1]);
        const uint64_t kh = umma_desc(sb + sh.off_k[0]), kl = umma_desc(sb + sh.off_k[1]);
        const uint32_t d = tmem_base + (uint32_t)b * kS1Col;
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_ss(d, ql + 2 * k, kh + 2 * k, sh.idesc_qk, (uint32_t)k);
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_ss(d, qh + 2 * k, kl + 2 * k, sh.idesc_qk, 1u);
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_ss(d, qh + 2 * k, kh + 2 * k, sh.idesc_qk, 1u);
        tc_commit(&bars[B_SFULL0 + b]);
        tc_commit(&bars[B_QFREE0 + b]);           // this query tile has been read
        if (b == 1) tc_commit(&bars[B_KFREE]);    // ... and so have the keys, by both tiles
      }
      __syncwarp();
    };
    auto issue_pv = [&](int b, uint32_t par, bool wait_v, bool free_v) {  // O = Pl Vh + Ph Vl + Ph Vh
      mbar_wait(&bars[B_PFULL0 + b], par);
      if (wait_v) mbar_wait(&bars[B_VFULL], par);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t vh = umma_desc(sb + sh.off_v[0]), vl = umma_desc(sb + sh.off_v[1]);
        const uint32_t pbase = tmem_base + (uint32_t)b * kS1Col, d = tmem_base + kOCol;
#pragma unroll
        for (uint32_t j = 0; j < nks; ++j) {  // 16 keys per instruction: 8 TMEM columns of P, 2 KB of V
          const bool tail = (NK % 32) != 0 && j == nks - 1;
          const uint32_t phi = pbase + 32 * (j >> 1) + (tail ? 0 : 8 * (j & 1));
          const uint32_t plo = phi + (tail ? 8 : 16);
          mma_ts(d, plo, vh + 128 * j, sh.idesc_pv, j);
          mma_ts(d, phi, vl + 128 * j, sh.idesc_pv, 1u);
          mma_ts(d, phi, vh + 128 * j, sh.idesc_pv, 1u);
        }
        tc_commit(&bars[B_OFULL0 + b]);
        if (free_v) tc_commit(&bars[B_VFREE]);  // both tiles' P V have read this item's values
      }
      __syncwarp();
    };
    int it = 0;
    for (int64_t item = blockIdx.x; item < sh.items; item += gridDim.x, ++it) {
      const uint32_t par = (uint32_t)it & 1;
      issue_qk(0, par);
      ATT_TRACE(8);
      if (it > 0) issue_pv(1, par ^ 1, false, true);
      ATT_TRACE(9);
      issue_qk(1, par);
      ATT_TRACE(10);
      issue_pv(0, par, true, false);
      ATT_TRACE(11);
    }
    if (it > 0) {
      // the last O1 claims the shared O columns: group 0 must have drained the last O0 (inside the loop the next
      // item's S0 waits for exactly that)
      mbar_wait(&bars[B_SFREE0], (uint32_t)(it - 1) & 1);
      issue_pv(1, (uint32_t)(it - 1) & 1, false, true);
    }
  } else {  // ===== softmax groups =====
    const int grp = warp >> 2;         // query tile / TMEM S buffer / staging buffer of this group
    const int quarter = warp & 3;      // TMEM lanes 32 * quarter .. + 31
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const uint32_t tbuf = tmem_base + lane_base + (uint32_t)grp * kS1Col;
    const uint32_t tob = tmem_base + lane_base + kOCol;
    uint8_t* ost = smem + sh.off_o[grp] + (size_t)quarter * 8192;  // this warp's 32 rows: hi tile | second / third tiles
    const bool warp_valid = grp * 128 + quarter * 32 < sh.T;
    const int Tn = sh.T;
    const float sl2 = sh.sl2;
    int it = 0;
    for (int64_t item = blockIdx.x; item < sh.items; item += gridDim.x, ++it) {
      const uint32_t par = (uint32_t)it & 1;
      mbar_wait(&bars[B_SFULL0 + grp], par);
      tc_fence_after();
      if (quarter == 0) ATT_TRACE(4 * grp + 0);
      float inv = 0.f;
      if (warp_valid) inv = softmax_row<NK>(tbuf, Tn, sl2);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_PFULL0 + grp]);
      if (quarter == 0) ATT_TRACE(4 * grp + 1);

      // ---- O = P V is on its way: make the staging tiles reusable meanwhile ----
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();

      mbar_wait(&bars[B_OFULL0 + grp], par);
      tc_fence_after();
      if (quarter == 0) ATT_TRACE(4 * grp + 2);
      uint32_t o[64];
      {
        uint32_t* o0 = o;
        uint32_t* o1 = o + 32;
        TMEM_LD_X32(tob, o0);
        TMEM_LD_X32(tob + 32, o1);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_SFREE0 + grp]);  // S and O buffers may be claimed again

      if (warp_valid) {
        // hi tile: rows of 128 bytes, 16-byte chunk j of row r at chunk (j ^ (r & 7)) = TMA SWIZZLE_128B
        uint8_t* hrow = ost + lane * 128;
        const int sw = lane & 7;
        if constexpr (OFMT == SVIT_FMT_X3) {
          uint8_t* lrow = ost + 4096 + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4 qh, ql;
            split_x3(__uint_as_float(o[8 * j]) * inv, __uint_as_float(o[8 * j + 1]) * inv, qh.x, ql.x);
            split_x3(__uint_as_float(o[8 * j + 2]) * inv, __uint_as_float(o[8 * j + 3]) * inv, qh.y, ql.y);
            split_x3(__uint_as_float(o[8 * j + 4]) * inv, __uint_as_float(o[8 * j + 5]) * inv, qh.z, ql.z);
            split_x3(__uint_as_float(o[8 * j + 6]) * inv, __uint_as_float(o[8 * j + 7]) * inv, qh.w, ql.w);
            *reinterpret_cast<uint4*>(hrow + ((j ^ sw) << 4)) = qh;
            *reinterpret_cast<uint4*>(lrow + ((j ^ sw) << 4)) = ql;
          }
        } else {
          // hi8 / lo8 tiles: rows of 64 bytes, 16-byte chunk j of row r at chunk (j ^ ((r >> 1) & 3)) = TMA SWIZZLE_64B
          uint8_t* arow = ost + 4096 + lane * 64;
          uint8_t* brow = ost + 6144 + lane * 64;
          const int s8 = (lane >> 1) & 3;
#pragma unroll
          for (int j = 0; j < 4; ++j) {  // 16 values: two hi16 chunks, one hi8 and one lo8 chunk
            uint32_t hw[8];
            uint16_t a8[8], b8[8];
#pragma unroll
            for (int q = 0; q < 8; ++q)
              split_c8(__uint_as_float(o[16 * j + 2 * q]) * inv, __uint_as_float(o[16 * j + 2 * q + 1]) * inv, hw[q], a8[q], b8[q]);
            *reinterpret_cast<uint4*>(hrow + (((2 * j) ^ sw) << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            *reinterpret_cast<uint4*>(hrow + (((2 * j + 1) ^ sw) << 4)) = make_uint4(hw[4], hw[5], hw[6], hw[7]);
            *reinterpret_cast<uint4*>(arow + ((j ^ s8) << 4)) =
                make_uint4((uint32_t)a8[0] | ((uint32_t)a8[1] << 16), (uint32_t)a8[2] | ((uint32_t)a8[3] << 16),
                           (uint32_t)a8[4] | ((uint32_t)a8[5] << 16), (uint32_t)a8[6] | ((uint32_t)a8[7] << 16));
            *reinterpret_cast<uint4*>(brow + ((j ^ s8) << 4)) =
                make_uint4((uint32_t)b8[0] | ((uint32_t)b8[1] << 16), (uint32_t)b8[2] | ((uint32_t)b8[3] << 16),
                           (uint32_t)b8[4] | ((uint32_t)b8[5] << 16), (uint32_t)b8[6] | ((uint32_t)b8[7] << 16));
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0 && warp_valid) {
        const int seq = (int)(item / sh.heads), head = (int)(item % sh.heads);
        const int row0 = grp * 128 + quarter * 32;  // 32-row boxes; rows >= T are clipped by the tensor maps
        tma_store_3d(&maps.o[0], ost, head * kD, row0, seq);
        if (OFMT == SVIT_FMT_X3) {
          tma_store_3d(&maps.o[1], ost + 4096, head * kD, row0, seq);
        } else {
          tma_store_3d(&maps.o[1], ost + 4096, head * kD, row0, seq);
          tma_store_3d(&maps.o[2], ost + 6144, head * kD, row0, seq);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      if (quarter == 0) ATT_TRACE(4 * grp + 3);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

struct SmemPlan {
  AttShape sh;
  size_t bytes;
};

// shared-memory carve-up for T tokens with the key bucket NK (every tile 1 KB aligned)
SmemPlan plan_smem(int Tn, int NK) {
  SmemPlan p{};
  const int R1 = (Tn - 128 + 7) / 8 * 8;
  uint32_t off = 0;
  auto take = [&](uint32_t bytes) {
    const uint32_t o = off;
    off += (bytes + 1023u) & ~1023u;
    return o;
  };
  p.sh.R1 = R1;
  p.sh.off_q[0][0] = take(128 * 128), p.sh.off_q[0][1] = take(128 * 128);
  p.sh.off_q[1][0] = take((uint32_t)R1 * 128), p.sh.off_q[1][1] = take((uint32_t)R1 * 128);
  p.sh.off_k[0] = take((uint32_t)NK * 128), p.sh.off_k[1] = take((uint32_t)NK * 128);
  p.sh.off_v[0] = take((uint32_t)NK * 128), p.sh.off_v[1] = take((uint32_t)NK * 128);
  p.sh.off_o[0] = take(4 * 8192);
  p.sh.off_o[1] = take((uint32_t)((R1 + 31) / 32) * 8192);
  // the second tile's MMAs read 128 rows from off_q[1][*]: rows >= R1 are whatever follows (their scores are never
  // stored), but they must lie inside the allocation: the staging tiles behind them guarantee that
  p.sh.off_bar = take(B_COUNT * 8 + 16);
  p.bytes = (size_t)off + 1024;  // + alignment slack
  return p;
}

int key_bucket(int Tn) { return Tn <= 160 ? 160 : Tn <= 208 ? 208 : 224; }

template <int NK, int OFMT>
int launch(const AttMaps& maps, const AttShape& sh, size_t smem, cudaStream_t stream) {
  auto kern = attention_tc_split_kernel<NK, OFMT>;
  SVIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t grid = std::min<int64_t>(sh.items, sm_count());
  kern<<<(unsigned)grid, kThreads, smem, stream>>>(maps, sh);
  SVIT_LAUNCH_CHECK("attention_tc_split_kernel");
  return SVIT_OK;
}

template <int OFMT>
int launch_nk(int NK, const AttMaps& maps, const AttShape& sh, size_t smem, cudaStream_t stream) {
  switch (NK) {
    case 160: return launch<160, OFMT>(maps, sh, smem, stream);
    case 208: return launch<208, OFMT>(maps, sh, smem, stream);
    default: return launch<224, OFMT>(maps, sh, smem, stream);
  }
}

}  // namespace

bool attention_split_tc_fits(int Tn) {
  if (Tn <= 128 || Tn > 224) return false;
  return plan_smem(Tn, key_bucket(Tn)).bytes <= kSmemLimit;
}

// qkv: X3 planes [n_seq, T, 3h]; ctx: X3 or C8 planes [n_seq, T, h]; head_dim 64, 128 < T <= 224
int attention_split_tc(const Operand& qkv, const Operand& ctx, int64_t n_seq, int Tn, int heads, cudaStream_t stream) {
  if (n_seq == 0) return SVIT_OK;
  SVIT_CHECK_ARG(attention_split_tc_fits(Tn), "attention_split_tc: T=%d out of range", Tn);
  const int h = heads * kD;
  const int NK = key_bucket(Tn);
  SmemPlan pl = plan_smem(Tn, NK);
  AttShape& sh = pl.sh;
  AttMaps maps;
  int rc;
  for (int p = 0; p < 2; ++p) {  // qkv [n_seq][T][3h]: boxes of 64 columns (one head of q, k or v) x rows
    const void* base = qkv.plane(p, 0);
    if ((rc = encode_map_3d(&maps.q0[p], SVIT_F16, base, 3 * h, Tn, n_seq, (uint64_t)3 * h * 2, (uint64_t)Tn * 3 * h * 2, kD, 128))) return rc;
    if ((rc = encode_map_3d(&maps.q1[p], SVIT_F16, base, 3 * h, Tn, n_seq, (uint64_t)3 * h * 2, (uint64_t)Tn * 3 * h * 2, kD, sh.R1))) return rc;
    if ((rc = encode_map_3d(&maps.kv[p], SVIT_F16, base, 3 * h, Tn, n_seq, (uint64_t)3 * h * 2, (uint64_t)Tn * 3 * h * 2, kD, NK))) return rc;
  }
  if ((rc = encode_map_3d(&maps.o[0], SVIT_F16, ctx.plane(0, 0), h, Tn, n_seq, (uint64_t)h * 2, (uint64_t)Tn * h * 2, kD, 32))) return rc;
  maps.o[2] = maps.o[0];
  if (ctx.fmt == SVIT_FMT_X3) {
    if ((rc = encode_map_3d(&maps.o[1], SVIT_F16, ctx.plane(1, 0), h, Tn, n_seq, (uint64_t)h * 2, (uint64_t)Tn * h * 2, kD, 32))) return rc;
  } else {
    if ((rc = encode_map_3d(&maps.o[1], SVIT_U8, ctx.plane(1, 0), h, Tn, n_seq, (uint64_t)h, (uint64_t)Tn * h, kD, 32, 64))) return rc;
    if ((rc = encode_map_3d(&maps.o[2], SVIT_U8, ctx.plane(2, 0), h, Tn, n_seq, (uint64_t)h, (uint64_t)Tn * h, kD, 32, 64))) return rc;
  }
  sh.T = Tn, sh.heads = heads;
  sh.items = n_seq * heads;
  sh.sl2 = 0.125f * 1.4426950408889634f;
  // D fp32 | fp16 A, B | (B MN-major for P V) | N >> 3 | M >> 4
  sh.idesc_qk = (1u << 4) | ((uint32_t)(NK >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  sh.idesc_pv = (1u << 4) | (1u << 16) | ((uint32_t)(kD >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  return ctx.fmt == SVIT_FMT_X3 ? launch_nk<SVIT_FMT_X3>(NK, maps, sh, pl.bytes, stream)
                                : launch_nk<SVIT_FMT_C8>(NK, maps, sh, pl.bytes, stream);
}

}  // namespace svit

// debug: device buffer of 24 x 16 int64 (or NULL to switch the trace off); not part of the public ABI
extern "C" int svit_debug_attention_split_trace(void* device_buffer) {
  using namespace svit;
  long long* p = static_cast<long long*>(device_buffer);
  SVIT_CUDA(cudaMemcpyToSymbol(g_att_split_trace, &p, sizeof(p)));
  return SVIT_OK;
}
