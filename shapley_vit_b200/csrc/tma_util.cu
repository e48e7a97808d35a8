// cuTensorMapEncodeTiled through the runtime's driver entry point (libsvit does not link libcuda).
#include "tma_util.h"

namespace svit {
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      return reinterpret_cast<EncodeTiledFn>(p);
    cudaGetLastError();
    return static_cast<EncodeTiledFn>(nullptr);
  }();
  return fn;
}

}  // namespace

int encode_map_3d(CUtensorMap* map, int dtype, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                  uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0, uint32_t box1, int swizzle_bytes) {
  EncodeTiledFn enc = get_encode();
  if (!enc) SVIT_FAIL(SVIT_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const CUtensorMapDataType dt = dtype == SVIT_F32    ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : dtype == SVIT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                 : dtype == SVIT_U8   ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                                      : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  cuuint64_t gdim[3] = {d0, d1, d2};
  cuuint64_t gstr[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapSwizzle sw = swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  const CUresult r = enc(map, dt, 3, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) SVIT_FAIL(SVIT_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return SVIT_OK;
}

}  // namespace svit
