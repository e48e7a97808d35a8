// Library plumbing (errors, device info) and the plan layout table.
//
// The plan layout is a fixed permutation of the HF ViT state_dict the reference aggregates
// key by key (federated_learning/utils.py:745-748, 787-791): a vec region (fp32) followed by
// a mat region (GEMM operand dtype).  shapley_vit_b200/layout.py computes the same table; a
// test asserts they agree.
#include "layout.h"

#include <cstring>
#include <mutex>

namespace svit {

namespace {
thread_local char g_err[512] = "";
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

int sm_count() {
  static int cached = -1;
  if (cached < 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

int validate_cfg(const svit_vit_cfg* c) {
  SVIT_CHECK_ARG(c != nullptr, "cfg is null");
  SVIT_CHECK_ARG(c->hidden > 0 && c->layers > 0 && c->heads > 0 && c->ff > 0 && c->image > 0 && c->patch > 0 &&
                     c->channels > 0 && c->n_cls > 0,
                 "cfg: all sizes must be positive");
  SVIT_CHECK_ARG(c->hidden % c->heads == 0, "cfg: hidden %% heads != 0");
  SVIT_CHECK_ARG(c->image % c->patch == 0, "cfg: image %% patch != 0");
  SVIT_CHECK_ARG(c->hidden % 64 == 0 && c->ff % 64 == 0, "cfg: hidden and ff must be multiples of 64");
  SVIT_CHECK_ARG((c->channels * c->patch * c->patch) % 64 == 0, "cfg: channels*patch*patch must be a multiple of 64");
  return SVIT_OK;
}

int build_layout(const svit_vit_cfg* c, Layout* L) {
  int rc = validate_cfg(c);
  if (rc) return rc;
  L->cfg = *c;
  L->segs.clear();
  const int64_t h = c->hidden, ff = c->ff, np = (int64_t)(c->image / c->patch) * (c->image / c->patch);
  const int64_t T = np + 1, pd = (int64_t)c->channels * c->patch * c->patch;
  int64_t off[2] = {0, 0};
  auto add = [&](int kind, int layer, int region, int64_t rows, int64_t cols) {
    svit_segment s{};
    s.kind = kind;
    s.layer = layer;
    s.region = region;
    s.offset = off[region];
    s.rows = rows;
    s.cols = cols;
    s.size = rows * cols;
    L->segs.push_back(s);
    off[region] += round_up(s.size, 64);
  };
  add(SVIT_SEG_CLS, -1, SVIT_REGION_VEC, 1, h);
  add(SVIT_SEG_POS, -1, SVIT_REGION_VEC, T, h);
  add(SVIT_SEG_PATCH_B, -1, SVIT_REGION_VEC, 1, h);
  for (int i = 0; i < c->layers; ++i) {
    add(SVIT_SEG_LN1_G, i, SVIT_REGION_VEC, 1, h);
    add(SVIT_SEG_LN1_B, i, SVIT_REGION_VEC, 1, h);
    add(SVIT_SEG_BQ, i, SVIT_REGION_VEC, 1, h);
    add(SVIT_SEG_BK, i, SVIT_REGION_VEC, 1, h);
    add(SVIT_SEG_BV, i, SVIT_REGION_VEC, 1, h);
    add(SVIT_SEG_BO, i, SVIT_REGION_VEC, 1, h);
    add(SVIT_SEG_LN2_G, i, SVIT_REGION_VEC, 1, h);
    add(SVIT_SEG_LN2_B, i, SVIT_REGION_VEC, 1, h);
    add(SVIT_SEG_B1, i, SVIT_REGION_VEC, 1, ff);
    add(SVIT_SEG_B2, i, SVIT_REGION_VEC, 1, h);
  }
  add(SVIT_SEG_LNF_G, -1, SVIT_REGION_VEC, 1, h);
  add(SVIT_SEG_LNF_B, -1, SVIT_REGION_VEC, 1, h);
  add(SVIT_SEG_HEAD_W, -1, SVIT_REGION_VEC, c->n_cls, h);
  add(SVIT_SEG_HEAD_B, -1, SVIT_REGION_VEC, 1, c->n_cls);
  add(SVIT_SEG_PATCH_W, -1, SVIT_REGION_MAT, h, pd);
  for (int i = 0; i < c->layers; ++i) {
    add(SVIT_SEG_WQ, i, SVIT_REGION_MAT, h, h);
    add(SVIT_SEG_WK, i, SVIT_REGION_MAT, h, h);
    add(SVIT_SEG_WV, i, SVIT_REGION_MAT, h, h);
    add(SVIT_SEG_WO, i, SVIT_REGION_MAT, h, h);
    add(SVIT_SEG_W1, i, SVIT_REGION_MAT, ff, h);
    add(SVIT_SEG_W2, i, SVIT_REGION_MAT, h, ff);
  }
  L->vec_size = off[0];
  L->mat_size = off[1];
  return SVIT_OK;
}

int64_t Layout::find(int kind, int layer) const {
  for (const auto& s : segs)
    if (s.kind == kind && s.layer == layer) return s.offset;
  return -1;
}

}  // namespace svit

extern "C" const char* svit_version(void) { return "libsvit 0.1.0 (sm_100a, abi 1)"; }
extern "C" const char* svit_last_error(void) { return svit::last_error(); }

extern "C" int svit_device_info(int* sm, int* major, int* minor) {
  using namespace svit;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    SVIT_FAIL(SVIT_ERR_NO_DEVICE, "no CUDA device visible");
  }
  int dev = 0;
  SVIT_CUDA(cudaGetDevice(&dev));
  int a = 0, b = 0, c = 0;
  SVIT_CUDA(cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev));
  SVIT_CUDA(cudaDeviceGetAttribute(&b, cudaDevAttrComputeCapabilityMajor, dev));
  SVIT_CUDA(cudaDeviceGetAttribute(&c, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm) *sm = a;
  if (major) *major = b;
  if (minor) *minor = c;
  if (b != 10) SVIT_FAIL(SVIT_ERR_NO_DEVICE, "device is sm_%d%d; libsvit is built for sm_100a only", b, c);
  return SVIT_OK;
}

extern "C" int svit_layout_sizes(const svit_vit_cfg* cfg, int64_t* vec_size, int64_t* mat_size, int32_t* n_segments) {
  svit::Layout L;
  int rc = svit::build_layout(cfg, &L);
  if (rc) return rc;
  if (vec_size) *vec_size = L.vec_size;
  if (mat_size) *mat_size = L.mat_size;
  if (n_segments) *n_segments = (int32_t)L.segs.size();
  return SVIT_OK;
}

extern "C" int svit_layout_segment(const svit_vit_cfg* cfg, int32_t index, svit_segment* out) {
  using namespace svit;
  Layout L;
  int rc = build_layout(cfg, &L);
  if (rc) return rc;
  SVIT_CHECK_ARG(out && index >= 0 && index < (int32_t)L.segs.size(), "svit_layout_segment: index %d out of range", index);
  *out = L.segs[index];
  return SVIT_OK;
}
