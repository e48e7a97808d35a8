// Microbenchmark: tcgen05.ld / tcgen05.st throughput per SM (bytes per clock).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define LD32(taddr, r) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" \
  : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr) : "memory")
#define ST8(taddr, r) asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory")
template <int MODE>  // 0: ld x32, wait each; 1: 4 ld x32 in flight then wait; 2: st x8
__global__ void k(uint32_t* out, long long* cyc, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      uint32_t r[32];
      LD32(base + (it & 7) * 32, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc ^= r[0] ^ r[31];
    } else if (MODE == 1) {
      uint32_t a[32], b[32], c[32], d[32];
      LD32(base, a); LD32(base + 32, b); LD32(base + 64, c); LD32(base + 96, d);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc ^= a[0] ^ b[1] ^ c[2] ^ d[3];
    } else {
      uint32_t r[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = acc + i;
      ST8(base + (it & 31) * 8, r);
      if ((it & 15) == 15) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(slot), "r"(512u) : "memory");
}
template <int MODE>
void run(const char* name, int warps, double bytes_per_iter_per_warp) {
  uint32_t* out; long long* cyc; long long h;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 4000;
  k<MODE><<<148, warps * 32>>>(out, cyc, iters);
  k<MODE><<<148, warps * 32>>>(out, cyc, iters);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-34s warps=%2d: %7.1f bytes/clk/SM  (%s)\n", name, warps, bytes_per_iter_per_warp * warps * iters / h, cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {1, 4, 8, 16}) {
    run<0>("ld 32x32b.x32, wait each", w, 4096);
    run<1>("ld 32x32b.x32 x4 in flight", w, 4 * 4096);
    run<2>("st 32x32b.x8", w, 1024);
  }
  return 0;
}
