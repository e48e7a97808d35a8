// Fused GEMM epilogue shared by the CUDA-core and the tcgen05 GEMM kernels.
//
// One definition of "what happens to an accumulator" so both GEMM back ends are
// bit-comparable: + bias -> (exact-erf GELU) -> row remap -> + rowvec -> + residual -> cast.
// It covers every dense op of the HF ViT layer (modeling_vit.py:228-230 QKV, :265-268 output
// projection + residual, :287-312 MLP with GELU + residual) and the patch embedding with the
// position-embedding add and the [CLS] row gap (:100-128).
#pragma once
#include "common.cuh"

namespace svit {

struct EpiArgs {
  const float* bias;      // [N] (+ g * bias_gs) or null
  const float* rowvec;    // [rows_out, N] (+ g * rowvec_gs) or null
  const float* residual;  // fp32 [M_out, N] (+ g * residual_gs) or null; may alias out
  void* out;              // [M_out, N] (+ g * out_gs) in out_dtype; split formats: the main plane of a packed operand array
  int64_t bias_gs, rowvec_gs, residual_gs, out_gs;
  int32_t gelu, rows_in, rows_out, row_shift, out_dtype;
  int32_t M, N;
  int32_t out_fmt;        // svit_operand_format of `out` (split formats: out_dtype is SVIT_F16)
  int64_t out_alloc;      // plane pitch of `out` in elements (split formats)
  __host__ __device__ __forceinline__ Operand out_operand() const { return Operand{out, out_fmt, out_alloc}; }
};

__device__ __forceinline__ int64_t epi_out_row(const EpiArgs& e, int r) {
  return e.rows_in > 0 ? (int64_t)(r / e.rows_in) * e.rows_out + e.row_shift + (r % e.rows_in) : (int64_t)r;
}

// value-level epilogue for one element; `orow` from epi_out_row
__device__ __forceinline__ float epi_apply(const EpiArgs& e, int g, int64_t orow, int n, float acc) {
  float v = acc;
  if (e.bias) v += e.bias[(size_t)g * e.bias_gs + n];
  if (e.gelu) v = gelu_erf(v);
  if (e.rowvec) v += e.rowvec[(size_t)g * e.rowvec_gs + (size_t)(orow % e.rows_out) * e.N + n];
  if (e.residual) v += e.residual[(size_t)g * e.residual_gs + (size_t)orow * e.N + n];
  return v;
}

__device__ __forceinline__ void epi_store(const EpiArgs& e, int g, int64_t orow, int n, float v) {
  const size_t idx = (size_t)g * e.out_gs + (size_t)orow * e.N + n;
  if (e.out_fmt != SVIT_FMT_PLAIN) {  // element-wise (slow path): the planes of one value
    const Operand o = e.out_operand();
    uint32_t hi;
    if (e.out_fmt == SVIT_FMT_X3) {
      uint32_t lo;
      split_x3(v, 0.f, hi, lo);
      reinterpret_cast<uint16_t*>(o.aux1())[idx] = (uint16_t)lo;
    } else {
      uint16_t h8, l8;
      split_c8(v, 0.f, hi, h8, l8);
      reinterpret_cast<uint8_t*>(o.aux1())[idx] = (uint8_t)h8;
      reinterpret_cast<uint8_t*>(o.aux2())[idx] = (uint8_t)l8;
    }
    reinterpret_cast<uint16_t*>(o.base)[idx] = (uint16_t)hi;
    return;
  }
  if (e.out_dtype == SVIT_F32)
    reinterpret_cast<float*>(e.out)[idx] = v;
  else if (e.out_dtype == SVIT_BF16)
    reinterpret_cast<__nv_bfloat16*>(e.out)[idx] = Cvt<__nv_bfloat16>::from_f(v);
  else
    reinterpret_cast<__half*>(e.out)[idx] = Cvt<__half>::from_f(v);
}

inline EpiArgs make_epi(const svit_epilogue* epi, void* out, int64_t out_gs, int out_dtype, int M, int N) {
  EpiArgs e{};
  if (epi) {
    e.bias = epi->bias;
    e.bias_gs = epi->bias_gs;
    e.rowvec = epi->rowvec;
    e.rowvec_gs = epi->rowvec_gs;
    e.residual = epi->residual;
    e.residual_gs = epi->residual_gs;
    e.gelu = epi->gelu;
    e.rows_in = epi->rows_in;
    e.rows_out = epi->rows_out;
    e.row_shift = epi->row_shift;
  }
  e.out = out;
  e.out_gs = out_gs;
  e.out_dtype = out_dtype;
  e.M = M;
  e.N = N;
  e.out_fmt = SVIT_FMT_PLAIN;
  e.out_alloc = 0;
  return e;
}
// `out` is a packed operand array (main plane at e.out) of `alloc` elements
inline void set_out_format(EpiArgs& e, int fmt, int64_t alloc) {
  e.out_fmt = fmt;
  e.out_alloc = alloc;
  if (fmt != SVIT_FMT_PLAIN) e.out_dtype = SVIT_F16;
}

// internal GEMM entry points (gemm_simt.cu / gemm_tc.cu)
int gemm_simt(int operand_dtype, const void* A, int64_t a_gs, const void* B, int64_t b_gs, int G, int M, int N, int K,
              const EpiArgs& epi, cudaStream_t stream);
// A / B: operand arrays in the format of `precision` (PLAIN for tf32 / bf16 / f16, X3 / C8 packed arrays for the
// split precisions), a_off / b_off the element offset of the operand inside its array
// K-extension of a grouped GEMM: out[g] = epilogue(A[g] B[g]^T + Ae[g] Be[g]^T) with Ae (M x 64) and Be (N x 64) in the
// operand format of the precision, both grouped.  Used for the low-rank per-coalition correction on top of a shared
// weight matrix (frozen-base LoRA): the correction rides the same accumulator as one more k-block per pass.
struct GemmExt {
  Operand A;
  int64_t a_off, a_gs;
  Operand B;
  int64_t b_off, b_gs;
};
constexpr int kGemmExtK = 64;

// b_gs == 0: B is shared by all groups (one [N, K] matrix).
int gemm_tc(int precision, const Operand& A, int64_t a_off, int64_t a_gs, const Operand& B, int64_t b_off, int64_t b_gs, int G,
            int M, int N, int K, const EpiArgs& epi, cudaStream_t stream, const GemmExt* ext = nullptr);

}  // namespace svit
