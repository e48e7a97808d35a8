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
inline int dtype_size(int dt) { return dt == SVIT_F32 ? 4 : 2; }

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
