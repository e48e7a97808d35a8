// Shared host/device helpers for libsvit (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/svit.h"

namespace svit {

// ---- error reporting (thread-local last error; no exceptions cross the ABI) -----------
void set_error(const char* fmt, ...);
const char* last_error();

#define SVIT_FAIL(code, ...)      \
  do {                            \
    ::svit::set_error(__VA_ARGS__); \
    return (code);                \
  } while (0)

#define SVIT_CHECK_ARG(cond, ...) \
  do {                            \
    if (!(cond)) SVIT_FAIL(SVIT_ERR_ARG, __VA_ARGS__); \
  } while (0)

#define SVIT_CUDA(call)                                                                     \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      SVIT_FAIL(SVIT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

#define SVIT_LAUNCH_CHECK(name)                                                             \
  do {                                                                                      \
    cudaError_t e__ = cudaGetLastError();                                                   \
    if (e__ != cudaSuccess)                                                                 \
      SVIT_FAIL(SVIT_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__));  \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int64_t round_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
constexpr int SVIT_U8 = 3;  // internal: one plane of 8-bit (e4m3) values, for tensor maps only
inline int dtype_size(int dt) { return dt == SVIT_F32 ? 4 : dt == SVIT_U8 ? 1 : 2; }

// ---- packed operand arrays (svit_operand_format, include/svit.h) ---------------------------------
// SVIT_FMT_C8: x*y ~ hi(x)*hi(y) + 2^-kC8ScaleD * (hi8(x)*lo8(y) + lo8(x)*hi8(y)), with
// hi8 = e4m3(kC8HiScale * hi), lo8 = e4m3(kC8LoScale * (x - hi)); kC8HiScale * kC8LoScale = 2^kC8ScaleD.
// The SAME scales for both roles, so any tensor can be the A or the B operand of a product.
// Ranges: hi8 saturates at |x| = 112, is normal down to |x| = 2^-8 (weights ~0.02 keep 3-4 bits);
// lo8 of an O(1) activation is ~2^0, of a 0.02 weight ~2^-5 (normal range starts at 2^-6).
#define SVIT_C8_HI_SCALE 4.0f
#define SVIT_C8_LO_SCALE 8192.0f
#define SVIT_C8_SCALE_D 15

struct Operand {
  void* base;     // main plane (PLAIN: the array itself)
  int fmt;        // svit_operand_format
  int64_t alloc;  // plane pitch in elements (split formats)
  __host__ __device__ __forceinline__ char* aux1() const { return static_cast<char*>(base) + 2 * alloc; }
  __host__ __device__ __forceinline__ char* aux2() const { return static_cast<char*>(base) + 3 * alloc; }
  // pointer to element `off` of plane k (0 = main, 1 / 2 = aux planes); es_plain: element size of a PLAIN array
  __host__ __device__ __forceinline__ char* plane(int k, int64_t off, int es_plain = 2) const {
    char* b = static_cast<char*>(base);
    if (fmt == SVIT_FMT_PLAIN) return b + off * es_plain;
    if (k == 0) return b + off * 2;
    if (fmt == SVIT_FMT_X3) return b + 2 * alloc + off * 2;
    return b + (k == 1 ? 2 : 3) * alloc + off;
  }
};
// element offset semantics: plane pointers for element index i (split formats: main is fp16)
//   X3: hi16 = (half*)base + i            lo16 = (half*)aux1() + i
//   C8: hi16 = (half*)base + i            hi8  = (u8*)aux1() + i          lo8 = (u8*)aux2() + i
inline Operand plain_operand(const void* p) { return Operand{const_cast<void*>(p), SVIT_FMT_PLAIN, 0}; }
inline int format_of_precision(int precision) {
  return precision == SVIT_PREC_F16X3 ? SVIT_FMT_X3 : precision == SVIT_PREC_F16C8 ? SVIT_FMT_C8 : SVIT_FMT_PLAIN;
}
// bytes of an operand array of `alloc` elements
inline int64_t operand_bytes(int fmt, int dtype, int64_t alloc) { return fmt == SVIT_FMT_PLAIN ? alloc * dtype_size(dtype) : alloc * 4; }

int sm_count();  // cached SM count of the current device (148 on B200)

// ---- device-side dtype helpers ---------------------------------------------------------
template <typename T> struct Cvt;
template <> struct Cvt<float> {
  static __device__ __forceinline__ float to_f(float v) { return v; }
  static __device__ __forceinline__ float from_f(float v) { return v; }
};
template <> struct Cvt<__nv_bfloat16> {
  static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};
template <> struct Cvt<__half> {
  static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
  // saturating (cvt.rn.satfinite): a finite fp32 value never becomes inf in the operand type
  static __device__ __forceinline__ __half from_f(float v) {
    unsigned short r;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(v));
    return __ushort_as_half(r);
  }
};

// two floats -> packed fp16x2 / bf16x2 (lo in the low half), one F2FP instruction
__device__ __forceinline__ uint32_t pack_f16x2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// two floats -> packed e4m3x2 (lo in the low byte), saturating
__device__ __forceinline__ uint16_t pack_e4m3x2(float lo, float hi) {
  uint16_t r;
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t h) {
  return __half22float2(*reinterpret_cast<const __half2*>(&h));
}
// x = hi + lo + O(2^-22 |x|): hi = fp16(x) (saturating), lo = fp16(x - hi) (x - hi is exact in fp32)
__device__ __forceinline__ void split_x3(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  hi = pack_f16x2_sat(v0, v1);
  const float2 h = unpack_f16x2(hi);
  lo = pack_f16x2_sat(__fsub_rn(v0, h.x), __fsub_rn(v1, h.y));
}
// C8 planes of two values: hi = fp16(x), hi8 = e4m3(4 hi), lo8 = e4m3(8192 (x - hi)).  hi8 is formed in fp16
// arithmetic straight from the packed hi (4 hi is exact in fp16; an overflow saturates to the same 448 as the fp32
// route): 8 instructions per pair instead of 11 -- this runs in every producer's inner loop.
__device__ __forceinline__ void split_c8(float v0, float v1, uint32_t& hi, uint16_t& hi8, uint16_t& lo8) {
  hi = pack_f16x2_sat(v0, v1);
  asm("{.reg .b32 t;\n\tmul.rn.f16x2 t, %1, %2;\n\tcvt.rn.satfinite.e4m3x2.f16x2 %0, t;}"
      : "=h"(hi8)
      : "r"(hi), "r"(0x44004400u));  // (4.0h, 4.0h)
  const float2 h = unpack_f16x2(hi);
  lo8 = pack_e4m3x2(__fsub_rn(v0, h.x) * SVIT_C8_LO_SCALE, __fsub_rn(v1, h.y) * SVIT_C8_LO_SCALE);
}
// 4 consecutive elements of a split-format array at element index i (i % 4 == 0)
template <int FMT>
__device__ __forceinline__ void store4_planes(const Operand& o, int64_t i, float a, float b, float c, float d) {
  static_assert(FMT == SVIT_FMT_X3 || FMT == SVIT_FMT_C8, "split formats only");
  if constexpr (FMT == SVIT_FMT_X3) {
    uint32_t h0, h1, l0, l1;
    split_x3(a, b, h0, l0);
    split_x3(c, d, h1, l1);
    *reinterpret_cast<uint2*>(static_cast<__half*>(o.base) + i) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(o.aux1()) + i) = make_uint2(l0, l1);
  } else {
    uint32_t h0, h1;
    uint16_t a0, a1, b0, b1;
    split_c8(a, b, h0, a0, b0);
    split_c8(c, d, h1, a1, b1);
    *reinterpret_cast<uint2*>(static_cast<__half*>(o.base) + i) = make_uint2(h0, h1);
    *reinterpret_cast<uint32_t*>(o.aux1() + i) = (uint32_t)a0 | ((uint32_t)a1 << 16);
    *reinterpret_cast<uint32_t*>(o.aux2() + i) = (uint32_t)b0 | ((uint32_t)b1 << 16);
  }
}
// value of element i of an X3 array
__device__ __forceinline__ float load_x3(const Operand& o, int64_t i) {
  return __half2float(static_cast<const __half*>(o.base)[i]) + __half2float(reinterpret_cast<const __half*>(o.aux1())[i]);
}
template <typename T> __device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) { return pack_f16x2_sat(lo, hi); }
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) { return pack_bf16x2(lo, hi); }

// 4 consecutive elements, one vector store (p must be aligned to 4 elements)
template <typename T>
__device__ __forceinline__ void store4(T* p, float a, float b, float c, float d);
template <>
__device__ __forceinline__ void store4<float>(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float a, float b, float c, float d) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
}
template <>
__device__ __forceinline__ void store4<__half>(__half* p, float a, float b, float c, float d) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_f16x2_sat(a, b), pack_f16x2_sat(c, d));
}

// exact-erf GELU, the HF "gelu" activation (torch.nn.functional.gelu, approximate='none')
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace svit
