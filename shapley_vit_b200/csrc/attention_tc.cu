// Fused-softmax attention on the 5th-generation tensor cores (tcgen05 + TMEM), 16-bit operands,
// head_dim 64, 128 < T <= 256 (ViT-B/L at 224 px: T = 197).
//
// Same contract as attention.cu / attention_mma.cu (HF ViTSelfAttention, modeling_vit.py:171-196,
// 220-252, as the reference evaluates it at federated_learning/utils.py:886): per (sequence, head)
//     ctx = softmax(q k^T / sqrt(d)) v,   non-causal, no mask.
//
// A persistent CTA per SM walks the (sequence, head) items.  Per item the query rows form two
// 128-row tiles (rows >= T are zero-filled by TMA and never stored); each tile owns one 256-column
// TMEM buffer and one group of four softmax warps (one per TMEM lane quarter):
//   warp 8      TMA producer: Q (2 boxes), K, V of the item -> 128B-swizzled smem, 2 stages
//   warp 9      MMA issuer:   S = Q K^T   (tcgen05.mma, A and B from smem, K-major, N = keys padded to 16)
//                             O = P V     (A = P read from TMEM, B = V from smem, MN-major)
//   warps 0..3  softmax group 0, warps 4..7 group 1: tcgen05.ld S (thread = query row) -> row max
//               -> exp2 with pre-scaled logits, fp32 row sum -> P as 16-bit pairs back into the SAME
//               TMEM columns (tcgen05.st) -> wait for O -> scale by 1 / sum -> smem -> TMA store.
// Scores and probabilities never leave the SM; the legacy mma.sync kernel this replaces ran the
// HMMA pipe at 43 % and was 20 % of the forward (profiles/r01_attention.md).
// The bound of this kernel is the MUFU (exp2) pipe: 16 results / clk / SM.
#include <cuda.h>

#include <algorithm>

#include "elementwise.h"
#include "tma_util.h"

namespace svit {
namespace {

constexpr int kD = 64;
constexpr int kThreads = 10 * 32;
// warps 0..7: softmax groups (group = warp / 4, TMEM lane quarter = warp % 4); the two single-lane roles get
// the HIGHEST warp ids: the issue arbiter prefers high warp ids, and the TMA / MMA issuers must never wait
// behind the arithmetic of the softmax warps they share a scheduler with
constexpr int kProducerWarp = 8, kMmaWarp = 9;
constexpr int kQBytes = 256 * 128;         // two 128-row query tiles, 128 bytes (64 x 16 bit) per row
constexpr int kKVBytes = 256 * 128;        // up to 256 keys
constexpr int kStageBytes = kQBytes + 2 * kKVBytes;
constexpr int kOBytes = 128 * 128;         // one 128-row output tile per softmax group
constexpr int kNumBars = 2 + 2 + 4 * 2;
constexpr size_t kSmem = 2 * (size_t)kStageBytes + 2 * kOBytes + kNumBars * 8 + 16 + 1024;
constexpr uint32_t kOCol = 128;            // O accumulator columns inside a tile's 256-column buffer

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
// packed fp32x2: two IEEE operations per issue slot (FFMA2 / FADD2)
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
// 2^x for two values x <= 0 on the FMA / ALU pipes instead of the MUFU pipe (16 results / clk / SM, the
// bound of this kernel): Cody-Waite split x = n + f, f in [-0.5, 0.5] via the 1.5 * 2^23 rounding
// trick, 2^f by a degree-4 polynomial (max relative error 2.7e-6, two decimal orders below the
// 16-bit rounding of P), 2^n by adding n to the exponent field.  Half of the exponentials of a row go
// this way, so both pipes work at the same time.
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  const float kMagic = 12582912.0f;  // 1.5 * 2^23
  x.x = fmaxf(x.x, -125.0f), x.y = fmaxf(x.y, -125.0f);
  const float2 t = fadd2(x, make_float2(kMagic, kMagic));
  const float2 f = fadd2(x, fadd2(make_float2(kMagic, kMagic), make_float2(-t.x, -t.y)));  // x - (t - magic)
  float2 p = make_float2(9.560510516e-03f, 9.560510516e-03f);
  p = ffma2(p, f, make_float2(5.591703951e-02f, 5.591703951e-02f));
  p = ffma2(p, f, make_float2(2.402498126e-01f, 2.402498126e-01f));
  p = ffma2(p, f, make_float2(6.931219697e-01f, 6.931219697e-01f));
  p = ffma2(p, f, make_float2(9.999991655e-01f, 9.999991655e-01f));
  return make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23)),
                     __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23)));
}
template <typename T> __device__ __forceinline__ uint32_t pack16(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack16<__half>(float lo, float hi) { return pack_f16x2_sat(lo, hi); }
template <> __device__ __forceinline__ uint32_t pack16<__nv_bfloat16>(float lo, float hi) { return pack_bf16x2(lo, hi); }

// Debug aid (scripts/attn_trace.py): when set, CTA 0 records clock64() per phase of its first items.
__device__ long long* g_att_trace = nullptr;
constexpr int kTraceItems = 24, kTraceSlots = 16;
#define ATT_TRACE(slot)                                                                      \
  do {                                                                                       \
    if (trace && lane == 0 && it < kTraceItems) trace[it * kTraceSlots + (slot)] = clock64(); \
  } while (0)

struct AttShape {
  int T, NK, heads;      // tokens, keys padded to a multiple of 16, heads
  int64_t items;         // n_seq * heads
  float sl2;             // d^-0.5 * log2(e)
  uint32_t idesc_qk, idesc_pv;
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

// second pass on one loaded chunk: p = 2^(s * sl2 + noff) accumulated into `sum` and written back as
// 16-bit pairs over the first half of the chunk's own columns.  MASK: keys >= Tn (only possible in
// the last chunk) contribute 0.
template <typename T, bool W16, bool MASK>
__device__ __forceinline__ void chunk_exp(uint32_t tbuf, int c0, int Tn, float sl2, float noff, const uint32_t (&r)[32],
                                          float2 (&sum)[2]) {
  constexpr int W = W16 ? 16 : 32;
  uint32_t w[16];
  const float2 sc = make_float2(sl2, sl2), of = make_float2(noff, noff);
#pragma unroll
  for (int i = 0; i < W; i += 2) {
    const float2 x = ffma2(make_float2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), sc, of);
    float2 e;
    if ((i >> 1) % 3 == 2)  // one pair in three (measured: 1/2 270 us, 1/3 260 us, 1/4 262 us per 1 024 x 12 heads)
      e = ex2_poly2(x);  // FMA / ALU pipes
    else
      e = make_float2(ex2_approx(x.x), ex2_approx(x.y));  // MUFU pipe
    if (MASK) {
      if (c0 + i >= Tn) e.x = 0.f;
      if (c0 + i + 1 >= Tn) e.y = 0.f;
    }
    sum[(i >> 1) & 1] = fadd2(sum[(i >> 1) & 1], e);
    w[i >> 1] = pack16<T>(e.x, e.y);
  }
  TMEM_ST_X8(tbuf + (uint32_t)(c0 >> 1), w);
  if (!W16) TMEM_ST_X8(tbuf + (uint32_t)(c0 >> 1) + 8, (w + 8));
}

// first pass on one loaded chunk: running row maxima (four independent chains: a single chain of
// dependent FMNMX3 was ~800 cycles per row, profiles/r08 trace)
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

// Softmax of the thread's row of S (NK columns at tbuf), P written in place; returns 1 / row sum.
// Both passes run their TMEM loads one chunk ahead of the arithmetic (two register buffers), and the
// first load of pass 2 is issued under the last maximum of pass 1: the load latency (not the MUFU
// pipe) was what bounded the first version of this kernel.
template <typename T, int NK>
__device__ __forceinline__ float softmax_row(uint32_t tbuf, int Tn, float sl2, long long* mid_stamp) {
  constexpr int NFULL = NK / 32;           // 32-column chunks
  constexpr bool TAIL = (NK % 32) != 0;    // plus one 16-column chunk
  constexpr int NCH = NFULL + (TAIL ? 1 : 0);
  // only the LAST chunk can hold keys >= T (NK - 16 < T <= NK)
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
    } else {
      chunk_max<false, false>(c * 32, Tn, cur, mx);
    }
  }
  const float m = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
  if (mid_stamp) *mid_stamp = clock64();
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
        chunk_exp<T, true, true>(tbuf, c * 32, Tn, sl2, noff, cur, sum);
      else
        chunk_exp<T, false, true>(tbuf, c * 32, Tn, sl2, noff, cur, sum);
    } else {
      chunk_exp<T, false, false>(tbuf, c * 32, Tn, sl2, noff, cur, sum);
    }
  }
  tmem_st_wait();
  return 1.0f / ((sum[0].x + sum[0].y) + (sum[1].x + sum[1].y));
}

// NK: keys padded to a multiple of 16 (compile time: the softmax loops are fully unrolled)
template <typename T, int NK>
__global__ void __launch_bounds__(kThreads, 1)
    attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                        const __grid_constant__ CUtensorMap map_o, const AttShape sh) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* o_stage = smem + 2 * (size_t)kStageBytes;
  uint64_t* kv_full = reinterpret_cast<uint64_t*>(o_stage + 2 * kOBytes);
  uint64_t* kv_empty = kv_full + 2;
  uint64_t* s_full = kv_empty + 2;  // [2] MMA -> softmax group: S ready
  uint64_t* p_full = s_full + 2;    // [2] softmax group -> MMA: P written
  uint64_t* o_full = p_full + 2;    // [2] MMA -> softmax group: O ready
  uint64_t* s_free = o_full + 2;    // [2] softmax group -> MMA: buffer drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long* const trace = blockIdx.x == 0 ? g_att_trace : nullptr;

  if (warp == kProducerWarp && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_kv) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_o) : "memory");
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&s_free[i], 4);
    }
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

  if (warp == kProducerWarp) {  // ===== TMA producer =====
    int it = 0;
    for (int64_t item = blockIdx.x; item < sh.items; item += gridDim.x, ++it) {
      const int s = it & 1;
      mbar_wait(&kv_empty[s], ((it >> 1) & 1) ^ 1);
      if (elect_one()) {
        const int seq = (int)(item / sh.heads), head = (int)(item % sh.heads);
        const int h = sh.heads * kD;
        uint8_t* st = smem + (size_t)s * kStageBytes;
        mbar_expect_tx(&kv_full[s], (uint32_t)(kQBytes + 2 * NK * 128));
        tma_load_3d(st, &map_q, &kv_full[s], head * kD, 0, seq);
        tma_load_3d(st + 128 * 128, &map_q, &kv_full[s], head * kD, 128, seq);
        tma_load_3d(st + kQBytes, &map_kv, &kv_full[s], h + head * kD, 0, seq);
        tma_load_3d(st + kQBytes + kKVBytes, &map_kv, &kv_full[s], 2 * h + head * kD, 0, seq);
      }
      __syncwarp();
    }
  } else if (warp == kMmaWarp) {  // ===== MMA issuer =====
    // Issue order per item i:  S0(i), O1(i-1), S1(i), O0(i): the softmax groups run out of step, so one
    // group's P V, O read-out and next Q K^T overlap the other group's exponentials.
    constexpr uint32_t nks = (uint32_t)NK / 16;
    auto issue_qk = [&](uint32_t st, int b, uint32_t par) {  // S_b = Q_b K^T
      mbar_wait(&s_free[b], par ^ 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t qdesc = umma_desc(st + (uint32_t)b * 128 * 128), kdesc = umma_desc(st + kQBytes);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma_ss(tmem_base + (uint32_t)b * 256, qdesc + 2 * k, kdesc + 2 * k, sh.idesc_qk, (uint32_t)k);
        tc_commit(&s_full[b]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](uint32_t st, int b, uint32_t par, uint64_t* stage_done) {  // O_b = P_b V
      mbar_wait(&p_full[b], par);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t vdesc = umma_desc(st + kQBytes + kKVBytes);
#pragma unroll
        for (uint32_t j = 0; j < nks; ++j)  // 16 keys per instruction: 8 TMEM columns of P, 2 KB of V
          mma_ts(tmem_base + (uint32_t)b * 256 + kOCol, tmem_base + (uint32_t)b * 256 + 8 * j, vdesc + 128 * j, sh.idesc_pv,
                 j);
        tc_commit(&o_full[b]);
        if (stage_done) tc_commit(stage_done);  // every MMA reading this smem stage has been issued
      }
      __syncwarp();
    };
    int it = 0;
    uint32_t st_prev = 0;
    for (int64_t item = blockIdx.x; item < sh.items; item += gridDim.x, ++it) {
      const int s = it & 1;
      const uint32_t par = (uint32_t)it & 1;
      mbar_wait(&kv_full[s], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t st = smem_u32(smem + (size_t)s * kStageBytes);
      ATT_TRACE(8);
      issue_qk(st, 0, par);
      ATT_TRACE(9);
      if (it > 0) issue_pv(st_prev, 1, par ^ 1, &kv_empty[s ^ 1]);
      ATT_TRACE(10);
      issue_qk(st, 1, par);
      ATT_TRACE(11);
      issue_pv(st, 0, par, nullptr);
      ATT_TRACE(12);
      st_prev = st;
    }
    if (it > 0) issue_pv(st_prev, 1, (uint32_t)(it - 1) & 1, &kv_empty[(it - 1) & 1]);
  } else {  // ===== softmax groups =====
    const int grp = warp >> 2;         // tile / TMEM buffer / staging buffer of this group
    const int quarter = warp & 3;      // TMEM lanes 32 * quarter .. + 31
    const uint32_t tbuf = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)grp * 256;
    uint8_t* ost = o_stage + (size_t)grp * kOBytes;
    const int row_in_tile = quarter * 32 + lane;
    const bool warp_valid = grp * 128 + quarter * 32 < sh.T;
    const int Tn = sh.T;
    const float sl2 = sh.sl2;
    int it = 0;
    for (int64_t item = blockIdx.x; item < sh.items; item += gridDim.x, ++it) {
      const uint32_t par = (uint32_t)it & 1;
      mbar_wait(&s_full[grp], par);
      tc_fence_after();
      if (quarter == 0) ATT_TRACE(4 * grp + 0);
      float inv = 0.f;
      if (warp_valid)
        inv = softmax_row<T, NK>(tbuf, Tn, sl2,
                                 nullptr);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[grp]);
      if (quarter == 0) ATT_TRACE(4 * grp + 1);

      // ---- O = P V is on its way: make the staging tile reusable meanwhile ----
      // (each warp stores its own 32 rows, so nothing here needs a group-wide barrier)
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();

      mbar_wait(&o_full[grp], par);
      tc_fence_after();
      if (quarter == 0) ATT_TRACE(4 * grp + 2);
      uint32_t o[64];
      {
        uint32_t* o0 = o;
        uint32_t* o1 = o + 32;
        TMEM_LD_X32(tbuf + kOCol, o0);
        TMEM_LD_X32(tbuf + kOCol + 32, o1);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[grp]);  // the buffer may receive the next item's S

      if (warp_valid) {  // rows of 128 bytes, 16-byte chunk j of row r at chunk (j ^ (r & 7)): TMA SWIZZLE_128B
        uint8_t* orow = ost + row_in_tile * 128;
        const int sw = row_in_tile & 7;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 q;
          q.x = pack16<T>(__uint_as_float(o[8 * j]) * inv, __uint_as_float(o[8 * j + 1]) * inv);
          q.y = pack16<T>(__uint_as_float(o[8 * j + 2]) * inv, __uint_as_float(o[8 * j + 3]) * inv);
          q.z = pack16<T>(__uint_as_float(o[8 * j + 4]) * inv, __uint_as_float(o[8 * j + 5]) * inv);
          q.w = pack16<T>(__uint_as_float(o[8 * j + 6]) * inv, __uint_as_float(o[8 * j + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + ((j ^ sw) << 4)) = q;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0 && warp_valid) {
        const int seq = (int)(item / sh.heads), head = (int)(item % sh.heads);
        // 32-row box; rows >= T are clipped by the tensor map
        tma_store_3d(&map_o, ost + quarter * 4096, head * kD, grp * 128 + quarter * 32, seq);
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

template <typename T, int NK>
int launch_att_nk(const CUtensorMap& mq, const CUtensorMap& mkv, const CUtensorMap& mo, const AttShape& sh,
                  cudaStream_t stream) {
  auto kern = attention_tc_kernel<T, NK>;
  SVIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
  const int64_t grid = std::min<int64_t>(sh.items, sm_count());
  kern<<<(unsigned)grid, kThreads, kSmem, stream>>>(mq, mkv, mo, sh);
  SVIT_LAUNCH_CHECK("attention_tc_kernel");
  return SVIT_OK;
}

template <typename T>
int launch_att(const void* qkv, void* ctx, int dtype, int64_t n_seq, int Tn, int heads, cudaStream_t stream) {
  const int h = heads * kD;
  const int NK = (Tn + 15) / 16 * 16;
  CUtensorMap mq, mkv, mo;
  int rc;
  // qkv [n_seq][T][3h]: boxes of 64 columns (one head of q, k or v) x rows
  if ((rc = encode_map_3d(&mq, dtype, qkv, 3 * h, Tn, n_seq, (uint64_t)3 * h * 2, (uint64_t)Tn * 3 * h * 2, kD, 128))) return rc;
  if ((rc = encode_map_3d(&mkv, dtype, qkv, 3 * h, Tn, n_seq, (uint64_t)3 * h * 2, (uint64_t)Tn * 3 * h * 2, kD, NK))) return rc;
  if ((rc = encode_map_3d(&mo, dtype, ctx, h, Tn, n_seq, (uint64_t)h * 2, (uint64_t)Tn * h * 2, kD, 32))) return rc;
  AttShape sh{};
  sh.T = Tn, sh.NK = NK, sh.heads = heads;
  sh.items = n_seq * heads;
  sh.sl2 = 0.125f * 1.4426950408889634f;
  const uint32_t fmt = dtype == SVIT_BF16 ? 1u : 0u;
  // D fp32 | A, B formats | (B MN-major for P V) | N >> 3 | M >> 4
  sh.idesc_qk = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(NK >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  sh.idesc_pv = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 16) | ((uint32_t)(kD >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  switch (NK) {
    case 144: return launch_att_nk<T, 144>(mq, mkv, mo, sh, stream);
    case 160: return launch_att_nk<T, 160>(mq, mkv, mo, sh, stream);
    case 176: return launch_att_nk<T, 176>(mq, mkv, mo, sh, stream);
    case 192: return launch_att_nk<T, 192>(mq, mkv, mo, sh, stream);
    case 208: return launch_att_nk<T, 208>(mq, mkv, mo, sh, stream);
    case 224: return launch_att_nk<T, 224>(mq, mkv, mo, sh, stream);
    case 240: return launch_att_nk<T, 240>(mq, mkv, mo, sh, stream);
    default: return launch_att_nk<T, 256>(mq, mkv, mo, sh, stream);
  }
}

}  // namespace

// undocumented debug hook: device buffer of kTraceItems * kTraceSlots int64 (or NULL to switch tracing off)
extern "C" int svit_debug_attention_trace(void* device_buffer) {
  long long* p = static_cast<long long*>(device_buffer);
  SVIT_CUDA(cudaMemcpyToSymbol(g_att_trace, &p, sizeof(p)));
  return SVIT_OK;
}

namespace {
}

// 16-bit operands, head_dim 64, 128 < T <= 256, h % 8 == 0, 16-byte aligned qkv / ctx
int attention_tc(const void* qkv, void* ctx, int dtype, int64_t n_seq, int Tn, int heads, cudaStream_t stream) {
  if (n_seq == 0) return SVIT_OK;
  SVIT_CHECK_ARG(Tn > 128 && Tn <= 256, "attention_tc: T=%d out of range (129..256)", Tn);
  SVIT_CHECK_ARG(n_seq <= 0x7fffffff, "attention: too many sequences");
  if (!aligned16(qkv) || !aligned16(ctx)) SVIT_FAIL(SVIT_ERR_ALIGN, "attention: qkv/ctx must be 16-byte aligned");
  if (dtype == SVIT_F16) return launch_att<__half>(qkv, ctx, dtype, n_seq, Tn, heads, stream);
  if (dtype == SVIT_BF16) return launch_att<__nv_bfloat16>(qkv, ctx, dtype, n_seq, Tn, heads, stream);
  SVIT_FAIL(SVIT_ERR_ARG, "attention_tc: dtype %d is not a 16-bit type", dtype);
}

}  // namespace svit
