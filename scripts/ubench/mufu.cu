// Microbenchmark: throughput of MUFU.EX2 alone and mixed with the softmax's other instructions.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu mufu.cu && ./mufu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned pk(float a, float b) { unsigned r; asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = -0.001f * (threadIdx.x + i);
  unsigned acc = 0;
  float s = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float e = ex2(v[i]);
      if (MODE == 0) v[i] = e - 1.0f;            // keep a dependency so nothing is hoisted; 1 FADD per MUFU
      if (MODE == 1) { v[i] = e - 1.0f; if (i & 1) acc ^= pk(e, v[i - 1]); }
      if (MODE == 2) { s += e; v[i] = fmaf(e, -0.5f, -0.1f); if (i & 1) acc ^= pk(e, v[i - 1]); }
    }
  }
  long long t1 = clock64();
  float r = s;
#pragma unroll
  for (int i = 0; i < 16; ++i) r += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r + acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(const char* name, int warps) {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  k<MODE><<<148, warps * 32>>>(out, cyc, iters);
  k<MODE><<<148, warps * 32>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per_smsp = (double)warps / 4 * iters * 16;  // MUFU warp-instructions per SMSP
  printf("%-28s warps/SM=%2d: %.2f cycles per MUFU warp-instruction per SMSP\n", name, warps, h / per_smsp);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {4, 8, 16}) {
    run<0>("MUFU.EX2 + FADD", w);
    run<1>("MUFU.EX2 + FADD + 0.5 F2FP", w);
    run<2>("MUFU + FADD + FFMA + 0.5 F2FP", w);
  }
  return 0;
}
