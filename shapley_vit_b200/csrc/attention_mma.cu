// Fused-softmax attention on tensor cores for 16-bit operands (fp16 / bf16), head_dim 64.
//
// Same contract as attention.cu (HF ViTSelfAttention, modeling_vit.py:171-196, 220-252, as the
// reference evaluates it at federated_learning/utils.py:886): per (sequence, head)
//     ctx = softmax(q k^T / sqrt(d)) v,   non-causal, T <= 256 (ViT: 197 / 5).
// One CTA per (sequence, head): Q, K, V of the head are staged once into XOR-swizzled shared
// memory with cp.async; each warp owns 16 query rows at a time and keeps the whole score row
// block S[16, Tpad] in registers (mma.sync.m16n8k16, fp32 accumulate), does the softmax in fp32
// registers (exp2 with pre-scaled logits, quad shuffles for the row max / sum), re-uses the S
// accumulators as the A fragments of P V, and writes ctx.  Scores and probabilities never leave
// the register file.  Attention is ~4 % of the forward's FLOPs; the warp-level MMA keeps the odd
// sequence length (197 = 12.3 x 16) cheap to tile, which a 128-row tcgen05 tile would not.
#include "elementwise.h"

namespace svit {
namespace {

constexpr int kD = 64;          // head dim
constexpr int kWarps = 4;

__device__ __forceinline__ uint32_t smem_u32a(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
template <typename T> struct Mma;
template <> struct Mma<__half> {
  static __device__ __forceinline__ void run(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  static __device__ __forceinline__ uint32_t pack(float x, float y) { return pack_f16x2_sat(x, y); }
};
template <> struct Mma<__nv_bfloat16> {
  static __device__ __forceinline__ void run(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  static __device__ __forceinline__ uint32_t pack(float x, float y) { return pack_bf16x2(x, y); }
};

// element (row, col) of a [rows][64] 16-bit tile, 16-byte chunks XOR-swizzled by row
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

// KT = padded sequence length / 16
template <typename T, int KT>
__global__ void __launch_bounds__(kWarps * 32) attention_mma_kernel(const T* __restrict__ qkv, T* __restrict__ ctx, int Tn,
                                                                     int heads) {
  constexpr int TP = KT * 16;
  extern __shared__ __align__(128) unsigned char att_raw[];
  unsigned char* Qs = att_raw;
  unsigned char* Ks = Qs + TP * 128;
  unsigned char* Vs = Ks + TP * 128;
  const int h = heads * kD;
  const int64_t seq = blockIdx.x;
  const int head = blockIdx.y;
  const T* base = qkv + seq * (int64_t)Tn * 3 * h + head * kD;
  const int tid = threadIdx.x;

  // ---- stage Q, K, V (rows >= Tn are zero) ----
  for (int i = tid; i < TP * 8; i += kWarps * 32) {
    const int row = i >> 3, chunk = i & 7;
    const uint32_t off = tile_off(row, chunk);
    if (row < Tn) {
      const T* src = base + (size_t)row * 3 * h + chunk * 8;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32a(Qs + off)), "l"(src));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32a(Ks + off)), "l"(src + h));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32a(Vs + off)), "l"(src + 2 * h));
    } else {
      const uint4 z = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(Qs + off) = z;
      *reinterpret_cast<uint4*>(Ks + off) = z;
      *reinterpret_cast<uint4*>(Vs + off) = z;
    }
  }
  asm volatile("cp.async.commit_group;");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t q_base = smem_u32a(Qs), k_base = smem_u32a(Ks), v_base = smem_u32a(Vs);
  const float sl2 = 0.125f * 1.4426950408889634f;  // d^-0.5 * log2(e), d = 64
  const int lm = lane >> 3, lr = lane & 7;          // ldmatrix: matrix index / row inside it

  for (int rt = warp; rt < KT; rt += kWarps) {
    const int r0 = rt * 16;
    // ---- S = Q K^T for 16 query rows x TP keys ----
    float S[2 * KT][4];
#pragma unroll
    for (int nt = 0; nt < 2 * KT; ++nt) S[nt][0] = S[nt][1] = S[nt][2] = S[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < kD / 16; ++ks) {
      uint32_t a[4];
      ldsm_x4(a[0], a[1], a[2], a[3], q_base + tile_off(r0 + lr + (lm & 1) * 8, ks * 2 + (lm >> 1)));
#pragma unroll
      for (int np = 0; np < KT; ++np) {  // 16 keys per ldmatrix.x4
        uint32_t b0, b1, b2, b3;
        ldsm_x4(b0, b1, b2, b3, k_base + tile_off(np * 16 + lr + (lm >> 1) * 8, ks * 2 + (lm & 1)));
        Mma<T>::run(S[2 * np], a, b0, b1);
        Mma<T>::run(S[2 * np + 1], a, b2, b3);
      }
    }
    // ---- softmax over keys (rows g and g+8 of this tile) ----
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 2 * KT; ++nt) {
      const int c = nt * 8 + 2 * t;
      if (c >= Tn) S[nt][0] = S[nt][2] = -INFINITY;
      if (c + 1 >= Tn) S[nt][1] = S[nt][3] = -INFINITY;
      m0 = fmaxf(m0, fmaxf(S[nt][0], S[nt][1]));
      m1 = fmaxf(m1, fmaxf(S[nt][2], S[nt][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    const float o0 = m0 * sl2, o1 = m1 * sl2;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 2 * KT; ++nt) {
      S[nt][0] = exp2f(fmaf(S[nt][0], sl2, -o0));
      S[nt][1] = exp2f(fmaf(S[nt][1], sl2, -o0));
      S[nt][2] = exp2f(fmaf(S[nt][2], sl2, -o1));
      S[nt][3] = exp2f(fmaf(S[nt][3], sl2, -o1));
      s0 += S[nt][0] + S[nt][1];
      s1 += S[nt][2] + S[nt][3];
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
    s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    const float i0 = 1.0f / s0, i1 = 1.0f / s1;
    // ---- O = P V (P normalised before rounding to 16 bits) ----
    float O[kD / 8][4];
#pragma unroll
    for (int dt = 0; dt < kD / 8; ++dt) O[dt][0] = O[dt][1] = O[dt][2] = O[dt][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) {
      uint32_t a[4];
      a[0] = Mma<T>::pack(S[2 * kk][0] * i0, S[2 * kk][1] * i0);
      a[1] = Mma<T>::pack(S[2 * kk][2] * i1, S[2 * kk][3] * i1);
      a[2] = Mma<T>::pack(S[2 * kk + 1][0] * i0, S[2 * kk + 1][1] * i0);
      a[3] = Mma<T>::pack(S[2 * kk + 1][2] * i1, S[2 * kk + 1][3] * i1);
#pragma unroll
      for (int dp = 0; dp < kD / 16; ++dp) {  // 16 channels per ldmatrix.x4.trans
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(b0, b1, b2, b3, v_base + tile_off(kk * 16 + lr + (lm & 1) * 8, dp * 2 + (lm >> 1)));
        Mma<T>::run(O[2 * dp], a, b0, b1);
        Mma<T>::run(O[2 * dp + 1], a, b2, b3);
      }
    }
    // ---- store ctx rows r0+g and r0+g+8 ----
    const int ra = r0 + g, rb = r0 + g + 8;
    T* out = ctx + seq * (int64_t)Tn * h + head * kD + 2 * t;
#pragma unroll
    for (int dt = 0; dt < kD / 8; ++dt) {
      if (ra < Tn) *reinterpret_cast<uint32_t*>(out + (size_t)ra * h + dt * 8) = Mma<T>::pack(O[dt][0], O[dt][1]);
      if (rb < Tn) *reinterpret_cast<uint32_t*>(out + (size_t)rb * h + dt * 8) = Mma<T>::pack(O[dt][2], O[dt][3]);
    }
  }
}

template <typename T, int KT>
int launch_mma(const void* qkv, void* ctx, int64_t n_seq, int Tn, int heads, cudaStream_t stream) {
  const size_t smem = (size_t)3 * KT * 16 * 128;
  auto kern = attention_mma_kernel<T, KT>;
  SVIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)n_seq, heads);
  kern<<<grid, kWarps * 32, smem, stream>>>((const T*)qkv, (T*)ctx, Tn, heads);
  SVIT_LAUNCH_CHECK("attention_mma_kernel");
  return SVIT_OK;
}

template <typename T>
int dispatch_kt(const void* qkv, void* ctx, int64_t n_seq, int Tn, int heads, cudaStream_t stream) {
  if (Tn <= 16) return launch_mma<T, 1>(qkv, ctx, n_seq, Tn, heads, stream);
  if (Tn <= 64) return launch_mma<T, 4>(qkv, ctx, n_seq, Tn, heads, stream);
  if (Tn <= 208) return launch_mma<T, 13>(qkv, ctx, n_seq, Tn, heads, stream);
  return launch_mma<T, 16>(qkv, ctx, n_seq, Tn, heads, stream);
}

// ---- split-precision variant (SVIT_PREC_F16X3 / F16C8): X3 planes in, split-format planes out ----
// Same tiling as attention_mma_kernel.  Q, K, V arrive as fp16 hi + lo planes (the QKV GEMM's epilogue
// emits them: x = hi + lo to ~2^-22); every product runs lo*hi + hi*lo + hi*hi into the fp32
// accumulators -- scores and P V both -- with P split in registers after the fp32 softmax, and the
// context goes out in the consumer GEMM's operand format (X3 or C8 planes).
// ~21 significant bits end to end at three MMAs per tile.  This is the warp-level (mma.sync) kernel for
// the shapes the tcgen05 kernel (attention_tc_split.cu) does not take: T <= 128 or T > 216.
// Only K and V are staged (hi + lo: 4 x 26 KB for T = 197), Q fragments come straight from global memory,
// so two CTAs of four warps fit an SM and cover each other's staging phase and tail.
constexpr int kWarps3 = 4;

__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) { split_x3(x, y, hi, lo); }

template <int FMT>
__device__ __forceinline__ void store2_planes(const Operand& o, int64_t i, float a, float b) {
  if constexpr (FMT == SVIT_FMT_X3) {
    uint32_t h, l;
    split_x3(a, b, h, l);
    *reinterpret_cast<uint32_t*>(static_cast<__half*>(o.base) + i) = h;
    *reinterpret_cast<uint32_t*>(reinterpret_cast<__half*>(o.aux1()) + i) = l;
  } else {
    uint32_t h;
    uint16_t h8, l8;
    split_c8(a, b, h, h8, l8);
    *reinterpret_cast<uint32_t*>(static_cast<__half*>(o.base) + i) = h;
    *reinterpret_cast<uint16_t*>(o.aux1() + i) = h8;
    *reinterpret_cast<uint16_t*>(o.aux2() + i) = l8;
  }
}

template <int KT, int OFMT>
__global__ void __launch_bounds__(kWarps3 * 32) attention_split_kernel(const Operand qkv, const Operand ctx, int Tn, int heads) {
  constexpr int TP = KT * 16;
  constexpr int TILE = TP * 128;
  extern __shared__ __align__(128) unsigned char att_raw[];
  unsigned char* Kh = att_raw;  // Kh | Kl | Vh | Vl
  const int h = heads * kD;
  const int64_t seq = blockIdx.x;
  const int head = blockIdx.y;
  const __half* qh = static_cast<const __half*>(qkv.base) + seq * (int64_t)Tn * 3 * h + head * kD;
  const __half* ql = reinterpret_cast<const __half*>(qkv.aux1()) + seq * (int64_t)Tn * 3 * h + head * kD;
  const int tid = threadIdx.x;

  for (int i = tid; i < TP * 8; i += kWarps3 * 32) {  // rows >= Tn are zero
    const int row = i >> 3, chunk = i & 7;
    const uint32_t off = tile_off(row, chunk);
    uint4 v[4] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
    if (row < Tn) {
      const size_t src = (size_t)row * 3 * h + chunk * 8;
      v[0] = __ldg(reinterpret_cast<const uint4*>(qh + src + h));
      v[1] = __ldg(reinterpret_cast<const uint4*>(ql + src + h));
      v[2] = __ldg(reinterpret_cast<const uint4*>(qh + src + 2 * h));
      v[3] = __ldg(reinterpret_cast<const uint4*>(ql + src + 2 * h));
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) *reinterpret_cast<uint4*>(Kh + m * TILE + off) = v[m];
  }
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t k_base = smem_u32a(Kh), v_base = k_base + 2 * TILE;
  const float sl2 = 0.125f * 1.4426950408889634f;  // d^-0.5 * log2(e), d = 64
  const int lm = lane >> 3, lr = lane & 7;
  using M = Mma<__half>;

  for (int rt = warp; rt < KT; rt += kWarps3) {
    const int r0 = rt * 16;
    float S[2 * KT][4];
#pragma unroll
    for (int nt = 0; nt < 2 * KT; ++nt) S[nt][0] = S[nt][1] = S[nt][2] = S[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < kD / 16; ++ks) {
      // A fragments of m16n8k16 straight from global memory: rows g / g + 8, columns 2t, 2t+1 (+8) of the k-block
      uint32_t ah[4] = {0, 0, 0, 0}, al[4] = {0, 0, 0, 0};
      {
        const size_t qa = (size_t)(r0 + g) * 3 * h + ks * 16 + 2 * t;
        const size_t qb = qa + (size_t)8 * 3 * h;
        if (r0 + g < Tn) {
          ah[0] = __ldg(reinterpret_cast<const uint32_t*>(qh + qa)), al[0] = __ldg(reinterpret_cast<const uint32_t*>(ql + qa));
          ah[2] = __ldg(reinterpret_cast<const uint32_t*>(qh + qa + 8)), al[2] = __ldg(reinterpret_cast<const uint32_t*>(ql + qa + 8));
        }
        if (r0 + g + 8 < Tn) {
          ah[1] = __ldg(reinterpret_cast<const uint32_t*>(qh + qb)), al[1] = __ldg(reinterpret_cast<const uint32_t*>(ql + qb));
          ah[3] = __ldg(reinterpret_cast<const uint32_t*>(qh + qb + 8)), al[3] = __ldg(reinterpret_cast<const uint32_t*>(ql + qb + 8));
        }
      }
#pragma unroll
      for (int np = 0; np < KT; ++np) {
        uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
        const uint32_t ko = tile_off(np * 16 + lr + (lm >> 1) * 8, ks * 2 + (lm & 1));
        ldsm_x4(h0, h1, h2, h3, k_base + ko);
        ldsm_x4(l0, l1, l2, l3, k_base + TILE + ko);
        M::run(S[2 * np], al, h0, h1);
        M::run(S[2 * np], ah, l0, l1);
        M::run(S[2 * np], ah, h0, h1);
        M::run(S[2 * np + 1], al, h2, h3);
        M::run(S[2 * np + 1], ah, l2, l3);
        M::run(S[2 * np + 1], ah, h2, h3);
      }
    }
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 2 * KT; ++nt) {
      const int c = nt * 8 + 2 * t;
      if (c >= Tn) S[nt][0] = S[nt][2] = -INFINITY;
      if (c + 1 >= Tn) S[nt][1] = S[nt][3] = -INFINITY;
      m0 = fmaxf(m0, fmaxf(S[nt][0], S[nt][1]));
      m1 = fmaxf(m1, fmaxf(S[nt][2], S[nt][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    const float o0 = m0 * sl2, o1 = m1 * sl2;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 2 * KT; ++nt) {
      S[nt][0] = exp2f(fmaf(S[nt][0], sl2, -o0));
      S[nt][1] = exp2f(fmaf(S[nt][1], sl2, -o0));
      S[nt][2] = exp2f(fmaf(S[nt][2], sl2, -o1));
      S[nt][3] = exp2f(fmaf(S[nt][3], sl2, -o1));
      s0 += S[nt][0] + S[nt][1];
      s1 += S[nt][2] + S[nt][3];
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
    s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    const float i0 = 1.0f / s0, i1 = 1.0f / s1;
    float O[kD / 8][4];
#pragma unroll
    for (int dt = 0; dt < kD / 8; ++dt) O[dt][0] = O[dt][1] = O[dt][2] = O[dt][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) {
      uint32_t ah[4], al[4];
      split2(S[2 * kk][0] * i0, S[2 * kk][1] * i0, ah[0], al[0]);
      split2(S[2 * kk][2] * i1, S[2 * kk][3] * i1, ah[1], al[1]);
      split2(S[2 * kk + 1][0] * i0, S[2 * kk + 1][1] * i0, ah[2], al[2]);
      split2(S[2 * kk + 1][2] * i1, S[2 * kk + 1][3] * i1, ah[3], al[3]);
#pragma unroll
      for (int dp = 0; dp < kD / 16; ++dp) {
        uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
        const uint32_t vo = tile_off(kk * 16 + lr + (lm & 1) * 8, dp * 2 + (lm >> 1));
        ldsm_x4_t(h0, h1, h2, h3, v_base + vo);
        ldsm_x4_t(l0, l1, l2, l3, v_base + TILE + vo);
        M::run(O[2 * dp], al, h0, h1);
        M::run(O[2 * dp], ah, l0, l1);
        M::run(O[2 * dp], ah, h0, h1);
        M::run(O[2 * dp + 1], al, h2, h3);
        M::run(O[2 * dp + 1], ah, l2, l3);
        M::run(O[2 * dp + 1], ah, h2, h3);
      }
    }
    const int ra = r0 + g, rb = r0 + g + 8;
    const int64_t out = seq * (int64_t)Tn * h + head * kD + 2 * t;
#pragma unroll
    for (int dt = 0; dt < kD / 8; ++dt) {
      if (ra < Tn) store2_planes<OFMT>(ctx, out + (int64_t)ra * h + dt * 8, O[dt][0], O[dt][1]);
      if (rb < Tn) store2_planes<OFMT>(ctx, out + (int64_t)rb * h + dt * 8, O[dt][2], O[dt][3]);
    }
  }
}

// ---- the [CLS] query alone (last encoder layer): one warp per (sequence, head), fp32 CUDA-core arithmetic on
// the reconstructed hi + lo values; ctx_cls [n_seq, h] in the consumer GEMM's split format.
// HBM-bound (every key and value row of the head is read once: 4 * T * 64 * 2 bytes per item): a quarter-warp
// covers one 128-byte row of a plane with 16-byte loads, four keys per warp iteration.
template <int OFMT>
__global__ void __launch_bounds__(256) attention_cls_split_kernel(const Operand qkv, const Operand ctx, int Tn, int heads,
                                                                  int64_t n_items) {
  __shared__ float p_s[8][256];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t item = blockIdx.x * 8LL + w;
  if (item >= n_items) return;
  const int64_t seq = item / heads;
  const int head = (int)(item % heads);
  const int h = heads * kD;
  const int g = lane >> 3, cb = lane & 7;  // key slot of the iteration, block of 8 channels
  const __half* hi = static_cast<const __half*>(qkv.base);
  const __half* lo = reinterpret_cast<const __half*>(qkv.aux1());
  const int64_t base = seq * (int64_t)Tn * 3 * h + head * kD + cb * 8;
  // 8 channels of one row, hi + lo -> fp32 (exact: 22 bits)
  auto load8 = [&](int64_t idx, float (&v)[8]) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(hi + idx));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(lo + idx));
    const uint32_t ah[4] = {a.x, a.y, a.z, a.w}, bl[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 x = unpack_f16x2(ah[i]), y = unpack_f16x2(bl[i]);
      v[2 * i] = x.x + y.x, v[2 * i + 1] = x.y + y.y;
    }
  };
  float q[8];
  load8(base, q);  // token 0 = [CLS]
  for (int j0 = 0; j0 < Tn; j0 += 4) {
    const int j = j0 + g;
    float dot = 0.f;
    if (j < Tn) {
      float k[8];
      load8(base + (int64_t)j * 3 * h + h, k);
#pragma unroll
      for (int i = 0; i < 8; ++i) dot = fmaf(q[i], k[i], dot);
    }
    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
    dot += __shfl_xor_sync(0xffffffffu, dot, 2);
    dot += __shfl_xor_sync(0xffffffffu, dot, 4);
    if (cb == 0 && j < Tn) p_s[w][j] = dot * 0.125f;
  }
  __syncwarp();
  float m = -INFINITY;
  for (int j = lane; j < Tn; j += 32) m = fmaxf(m, p_s[w][j]);
  m = warp_max(m);
  float sum = 0.f;
  for (int j = lane; j < Tn; j += 32) {
    const float e = expf(p_s[w][j] - m);
    p_s[w][j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  __syncwarp();
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = 0.f;
  for (int j0 = 0; j0 < Tn; j0 += 4) {
    const int j = j0 + g;
    if (j < Tn) {
      float v[8];
      load8(base + (int64_t)j * 3 * h + 2 * h, v);
      const float pj = p_s[w][j] * inv;
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaf(pj, v[i], o[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {  // the four key slots
    o[i] += __shfl_xor_sync(0xffffffffu, o[i], 8);
    o[i] += __shfl_xor_sync(0xffffffffu, o[i], 16);
  }
  if (g == 0) {
    const int64_t oi = seq * (int64_t)h + head * kD + cb * 8;
    store4_planes<OFMT>(ctx, oi, o[0], o[1], o[2], o[3]);
    store4_planes<OFMT>(ctx, oi + 4, o[4], o[5], o[6], o[7]);
  }
}

template <int KT, int OFMT>
int launch_split(const Operand& qkv, const Operand& ctx, int64_t n_seq, int Tn, int heads, cudaStream_t stream) {
  const size_t smem = (size_t)4 * KT * 16 * 128;
  auto kern = attention_split_kernel<KT, OFMT>;
  SVIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)n_seq, heads);
  kern<<<grid, kWarps3 * 32, smem, stream>>>(qkv, ctx, Tn, heads);
  SVIT_LAUNCH_CHECK("attention_split_kernel");
  return SVIT_OK;
}

template <int OFMT>
int dispatch_split(const Operand& qkv, const Operand& ctx, int64_t n_seq, int Tn, int heads, bool cls_only, cudaStream_t stream) {
  if (cls_only) {
    const int64_t items = n_seq * heads;
    attention_cls_split_kernel<OFMT><<<(unsigned)((items + 7) / 8), 256, 0, stream>>>(qkv, ctx, Tn, heads, items);
    SVIT_LAUNCH_CHECK("attention_cls_split_kernel");
    return SVIT_OK;
  }
  if (Tn <= 16) return launch_split<1, OFMT>(qkv, ctx, n_seq, Tn, heads, stream);
  if (Tn <= 64) return launch_split<4, OFMT>(qkv, ctx, n_seq, Tn, heads, stream);
  if (Tn <= 208) return launch_split<13, OFMT>(qkv, ctx, n_seq, Tn, heads, stream);
  return launch_split<16, OFMT>(qkv, ctx, n_seq, Tn, heads, stream);
}

}  // namespace

// 16-bit operands, head_dim 64, T <= 256; qkv/ctx must be 16-byte aligned with h % 8 == 0
int attention_mma(const void* qkv, void* ctx, int dtype, int64_t n_seq, int Tn, int heads, cudaStream_t stream) {
  if (n_seq == 0) return SVIT_OK;
  SVIT_CHECK_ARG(Tn >= 1 && Tn <= 256, "attention: T=%d out of range (1..256)", Tn);
  SVIT_CHECK_ARG(n_seq <= 0x7fffffff, "attention: too many sequences");
  if (!aligned16(qkv) || !aligned16(ctx)) SVIT_FAIL(SVIT_ERR_ALIGN, "attention: qkv/ctx must be 16-byte aligned");
  if (dtype == SVIT_F16) return dispatch_kt<__half>(qkv, ctx, n_seq, Tn, heads, stream);
  if (dtype == SVIT_BF16) return dispatch_kt<__nv_bfloat16>(qkv, ctx, n_seq, Tn, heads, stream);
  SVIT_FAIL(SVIT_ERR_ARG, "attention_mma: dtype %d is not a 16-bit type", dtype);
}

// X3 qkv planes, split-format ctx planes, head_dim 64, T <= 256: the attention of the split precisions
int attention_split(const Operand& qkv, const Operand& ctx, int64_t n_seq, int Tn, int heads, int head_dim, bool cls_only,
                    cudaStream_t stream) {
  if (n_seq == 0) return SVIT_OK;
  SVIT_CHECK_ARG(Tn >= 1 && Tn <= 256, "attention: T=%d out of range (1..256)", Tn);
  SVIT_CHECK_ARG(n_seq <= 0x7fffffff, "attention: too many sequences");
  if (head_dim != kD) SVIT_FAIL(SVIT_ERR_UNSUPPORTED, "attention_split: head_dim %d (the split precisions take head_dim 64)", head_dim);
  SVIT_CHECK_ARG(qkv.fmt == SVIT_FMT_X3 && (ctx.fmt == SVIT_FMT_X3 || ctx.fmt == SVIT_FMT_C8),
                 "attention_split: qkv must be X3 planes and ctx X3 or C8 planes");
  if (!aligned16(qkv.base) || !aligned16(ctx.base) || qkv.alloc % 16 || ctx.alloc % 16)
    SVIT_FAIL(SVIT_ERR_ALIGN, "attention: qkv/ctx must be 16-byte aligned with plane pitches multiples of 16");
  static const bool no_tc = [] {
    const char* e = getenv("SVIT_ATTENTION_MMA_SYNC");
    return e && e[0] == '1';
  }();
  if (!cls_only && !no_tc && Tn > 128 && attention_split_tc_fits(Tn)) return attention_split_tc(qkv, ctx, n_seq, Tn, heads, stream);
  return ctx.fmt == SVIT_FMT_X3 ? dispatch_split<SVIT_FMT_X3>(qkv, ctx, n_seq, Tn, heads, cls_only, stream)
                                : dispatch_split<SVIT_FMT_C8>(qkv, ctx, n_seq, Tn, heads, cls_only, stream);
}

}  // namespace svit
