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

// ---- split-precision variant (SVIT_PREC_F16X3): fp32 in, fp32 out -------------------------------
// Same tiling as attention_mma_kernel, but Q, K, V arrive in fp32 and are staged as fp16 hi + lo
// tiles (x = hi + lo to ~2^-22); every product runs hi*hi + hi*lo + lo*hi into the fp32
// accumulators -- scores and P V both -- with P split in registers after the fp32 softmax.
// ~21 significant bits end to end at three MMAs per tile: the attention of the f16x3 mode.
// Only K and V are staged (hi + lo: 4 x 26 KB for T = 197), Q fragments are read from global memory and
// split in registers, so two CTAs of four warps fit an SM and cover each other's staging phase and tail.
constexpr int kWarps3 = 4;

__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
  hi = pack_f16x2_sat(x, y);
  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  lo = pack_f16x2_sat(x - hf.x, y - hf.y);
}

__device__ __forceinline__ void split8(const float* __restrict__ src, unsigned char* hi_tile, unsigned char* lo_tile,
                                       uint32_t off) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
  uint4 hi, lo;
  split2(a.x, a.y, hi.x, lo.x);
  split2(a.z, a.w, hi.y, lo.y);
  split2(b.x, b.y, hi.z, lo.z);
  split2(b.z, b.w, hi.w, lo.w);
  *reinterpret_cast<uint4*>(hi_tile + off) = hi;
  *reinterpret_cast<uint4*>(lo_tile + off) = lo;
}

template <int KT>
__global__ void __launch_bounds__(kWarps3 * 32) attention_split_kernel(const float* __restrict__ qkv, float* __restrict__ ctx,
                                                                        int Tn, int heads) {
  constexpr int TP = KT * 16;
  constexpr int TILE = TP * 128;
  extern __shared__ __align__(128) unsigned char att_raw[];
  unsigned char* Kh = att_raw;  // Kh | Kl | Vh | Vl
  const int h = heads * kD;
  const int64_t seq = blockIdx.x;
  const int head = blockIdx.y;
  const float* base = qkv + seq * (int64_t)Tn * 3 * h + head * kD;
  const int tid = threadIdx.x;

  for (int i = tid; i < TP * 8; i += kWarps3 * 32) {  // rows >= Tn are zero
    const int row = i >> 3, chunk = i & 7;
    const uint32_t off = tile_off(row, chunk);
    if (row < Tn) {
      const float* src = base + (size_t)row * 3 * h + chunk * 8;
      split8(src + h, Kh, Kh + TILE, off);
      split8(src + 2 * h, Kh + 2 * TILE, Kh + 3 * TILE, off);
    } else {
      const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int m = 0; m < 4; ++m) *reinterpret_cast<uint4*>(Kh + m * TILE + off) = z;
    }
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
      uint32_t ah[4], al[4];
      {
        const float* qa = base + (size_t)(r0 + g) * 3 * h + ks * 16 + 2 * t;
        const float* qb = qa + (size_t)8 * 3 * h;
        const float2 z = make_float2(0.f, 0.f);
        const bool va = r0 + g < Tn, vb = r0 + g + 8 < Tn;
        const float2 q0 = va ? __ldg(reinterpret_cast<const float2*>(qa)) : z;
        const float2 q1 = vb ? __ldg(reinterpret_cast<const float2*>(qb)) : z;
        const float2 q2 = va ? __ldg(reinterpret_cast<const float2*>(qa + 8)) : z;
        const float2 q3 = vb ? __ldg(reinterpret_cast<const float2*>(qb + 8)) : z;
        split2(q0.x, q0.y, ah[0], al[0]);
        split2(q1.x, q1.y, ah[1], al[1]);
        split2(q2.x, q2.y, ah[2], al[2]);
        split2(q3.x, q3.y, ah[3], al[3]);
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
    float* out = ctx + seq * (int64_t)Tn * h + head * kD + 2 * t;
#pragma unroll
    for (int dt = 0; dt < kD / 8; ++dt) {
      if (ra < Tn) *reinterpret_cast<float2*>(out + (size_t)ra * h + dt * 8) = make_float2(O[dt][0], O[dt][1]);
      if (rb < Tn) *reinterpret_cast<float2*>(out + (size_t)rb * h + dt * 8) = make_float2(O[dt][2], O[dt][3]);
    }
  }
}

template <int KT>
int launch_split(const float* qkv, float* ctx, int64_t n_seq, int Tn, int heads, cudaStream_t stream) {
  const size_t smem = (size_t)4 * KT * 16 * 128;
  auto kern = attention_split_kernel<KT>;
  SVIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)n_seq, heads);
  kern<<<grid, kWarps3 * 32, smem, stream>>>(qkv, ctx, Tn, heads);
  SVIT_LAUNCH_CHECK("attention_split_kernel");
  return SVIT_OK;
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

// fp32 qkv / ctx, head_dim 64, T <= 256: split-precision tensor-core attention (SVIT_PREC_F16X3)
int attention_split(const float* qkv, float* ctx, int64_t n_seq, int Tn, int heads, cudaStream_t stream) {
  if (n_seq == 0) return SVIT_OK;
  SVIT_CHECK_ARG(Tn >= 1 && Tn <= 256, "attention: T=%d out of range (1..256)", Tn);
  SVIT_CHECK_ARG(n_seq <= 0x7fffffff, "attention: too many sequences");
  if (!aligned16(qkv) || !aligned16(ctx)) SVIT_FAIL(SVIT_ERR_ALIGN, "attention: qkv/ctx must be 16-byte aligned");
  if (Tn <= 16) return launch_split<1>(qkv, ctx, n_seq, Tn, heads, stream);
  if (Tn <= 64) return launch_split<4>(qkv, ctx, n_seq, Tn, heads, stream);
  if (Tn <= 208) return launch_split<13>(qkv, ctx, n_seq, Tn, heads, stream);
  return launch_split<16>(qkv, ctx, n_seq, Tn, heads, stream);
}

}  // namespace svit
