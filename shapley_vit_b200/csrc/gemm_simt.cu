// CUDA-core GEMM with fp32 FMA accumulation: the exact mode (SVIT_PREC_F32) and the
// bit-comparable cross-check for the tcgen05 kernels (same operands, same epilogue).
// out[g] = epilogue(A[g] (M x K) * B[g]^T (N x K)), both operands K-contiguous, like every
// nn.Linear of the HF ViT the reference evaluates (federated_learning/utils.py:886).
#include "epilogue.cuh"

namespace svit {
namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;

template <typename T>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const T* __restrict__ A, int64_t a_gs, const T* __restrict__ B,
                                                        int64_t b_gs, int M, int N, int K, const EpiArgs epi) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int g = blockIdx.z;
  const T* Ag = A + (size_t)g * a_gs;
  const T* Bg = B + (size_t)g * b_gs;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;  // 16 x 16 threads, each a 4 x 4 micro-tile
  float acc[TM][TN] = {};
  // loader mapping: 64 rows x 16 k = 1024 elements, 4 consecutive k per thread
  const int lr = tid / 4, lk = (tid % 4) * 4;
  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = k0 + lk + q;
      const int am = m0 + lr, bn = n0 + lr;
      As[lk + q][lr] = (am < M && k < K) ? Cvt<T>::to_f(Ag[(size_t)am * K + k]) : 0.f;
      Bs[lk + q][lr] = (bn < N && k < K) ? Cvt<T>::to_f(Bg[(size_t)bn * K + k]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int r = m0 + ty * TM + i;
    if (r >= M) continue;
    const int64_t orow = epi_out_row(epi, r);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n < N) epi_store(epi, g, orow, n, epi_apply(epi, g, orow, n, acc[i][j]));
    }
  }
}

}  // namespace

int gemm_simt(int operand_dtype, const void* A, int64_t a_gs, const void* B, int64_t b_gs, int G, int M, int N, int K,
              const EpiArgs& epi, cudaStream_t stream) {
  if (G <= 0 || M <= 0 || N <= 0) return SVIT_OK;
  SVIT_CHECK_ARG(G <= 65535, "gemm_simt: too many groups");
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, G);
  SVIT_CHECK_ARG(grid.y <= 65535, "gemm_simt: M too large for this back end");
  switch (operand_dtype) {
    case SVIT_F32:
      gemm_simt_kernel<float><<<grid, 256, 0, stream>>>((const float*)A, a_gs, (const float*)B, b_gs, M, N, K, epi);
      break;
    case SVIT_BF16:
      gemm_simt_kernel<__nv_bfloat16>
          <<<grid, 256, 0, stream>>>((const __nv_bfloat16*)A, a_gs, (const __nv_bfloat16*)B, b_gs, M, N, K, epi);
      break;
    case SVIT_F16:
      gemm_simt_kernel<__half><<<grid, 256, 0, stream>>>((const __half*)A, a_gs, (const __half*)B, b_gs, M, N, K, epi);
      break;
    default: SVIT_FAIL(SVIT_ERR_ARG, "gemm_simt: bad dtype %d", operand_dtype);
  }
  SVIT_LAUNCH_CHECK("gemm_simt_kernel");
  return SVIT_OK;
}

}  // namespace svit
