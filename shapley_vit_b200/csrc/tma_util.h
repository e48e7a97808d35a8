// TMA tensor-map encoding shared by the tcgen05 kernels (gemm_tc.cu, attention_tc.cu).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace svit {

// 3-D map over a row-major [d2][d1][d0] tensor of `dtype` (svit_dtype) elements, d0 contiguous;
// stride1_bytes / stride2_bytes are the byte pitches of d1 and d2.  Box = box0 x box1 x 1 elements,
// swizzle of swizzle_bytes (128, or 64; box0 * element size must equal it), out-of-bounds elements read as
// zero and are not written.  dtype: svit_dtype or SVIT_U8 (one plane of e4m3 values).
int encode_map_3d(CUtensorMap* map, int dtype, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                  uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0, uint32_t box1, int swizzle_bytes = 128);

}  // namespace svit
