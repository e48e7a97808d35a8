// K5 -- on-device scoring of a batch of coalitions.
//
// Replaces the per-batch tail of the reference's `evaluation`
// (federated_learning/utils.py:891-894):
//     pred = outputs.argmax(dim=1); correct += pred.eq(labels).sum().item()
//     loss += CrossEntropyLoss(reduction='sum')(outputs, labels).item()
// which costs two host synchronisations per 128-image batch per coalition.  Here one CTA per
// coalition walks all n samples; nothing returns to the host until the caller reads the
// [C] counters.  The reduction order is fixed (strided per-thread partials -> shuffle tree ->
// per-warp partials in order), so the result does not depend on how coalitions are split
// across launches or GPUs.
#include <cmath>

#include "common.cuh"

namespace svit {
namespace {

constexpr int kScoreBlock = 256;

__global__ void __launch_bounds__(kScoreBlock) score_kernel(const float* __restrict__ logits, int64_t logits_stride,
                                                            const int64_t* __restrict__ labels, int64_t n, int n_cls,
                                                            int64_t* __restrict__ correct,
                                                            double* __restrict__ loss_sum, int out_stride,
                                                            int32_t* __restrict__ pred, int64_t pred_stride,
                                                            int accumulate) {
  const int c = blockIdx.x;
  const float* lg = logits + (size_t)c * logits_stride;
  long long my_correct = 0;
  double my_loss = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += kScoreBlock) {
    const float* row = lg + i * n_cls;
    float m = row[0];
    int arg = 0;
    bool has_nan = isnan(m);
    for (int k = 1; k < n_cls; ++k) {
      const float v = row[k];
      if (isnan(v) && !has_nan) {  // torch.argmax: the first NaN wins
        has_nan = true;
        arg = k;
      }
      if (!has_nan && v > m) {  // strict '>' keeps the first maximal index
        m = v;
        arg = k;
      }
    }
    const int64_t lab = labels[i];
    float ce;
    if (has_nan || lab < 0 || lab >= n_cls) {
      ce = nanf("");
    } else {
      float s = 0.f;
      for (int k = 0; k < n_cls; ++k) s += expf(row[k] - m);
      ce = (logf(s) + m) - row[lab];
    }
    my_correct += (arg == (int)lab) ? 1 : 0;
    my_loss += (double)ce;
    if (pred) pred[(size_t)c * pred_stride + i] = arg;
  }
  // fixed-order block reduction
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    my_correct += __shfl_down_sync(0xffffffffu, my_correct, o);
    my_loss += __shfl_down_sync(0xffffffffu, my_loss, o);
  }
  __shared__ long long s_c[kScoreBlock / 32];
  __shared__ double s_l[kScoreBlock / 32];
  if ((threadIdx.x & 31) == 0) {
    s_c[threadIdx.x >> 5] = my_correct;
    s_l[threadIdx.x >> 5] = my_loss;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long tc = 0;
    double tl = 0.0;
    for (int w = 0; w < kScoreBlock / 32; ++w) {
      tc += s_c[w];
      tl += s_l[w];
    }
    const size_t o = (size_t)c * out_stride;  // 1: separate [C] arrays; 2: interleaved svit_record entries
    if (accumulate) {
      correct[o] += tc;
      loss_sum[o] += tl;
    } else {
      correct[o] = tc;
      loss_sum[o] = tl;
    }
  }
}

}  // namespace
}  // namespace svit

extern "C" int svit_score(const float* logits, int64_t logits_stride, const int64_t* labels, int C, int64_t n,
                          int n_cls, int64_t* correct, double* loss_sum, int32_t* pred, int64_t pred_stride,
                          int accumulate, svit_stream_t stream) {
  using namespace svit;
  SVIT_CHECK_ARG(logits && labels && correct && loss_sum, "svit_score: null pointer");
  SVIT_CHECK_ARG(C >= 1 && n >= 0 && n_cls >= 1, "svit_score: C=%d n=%lld n_cls=%d out of range", C, (long long)n, n_cls);
  SVIT_CHECK_ARG(logits_stride >= n * n_cls, "svit_score: logits_stride too small");
  SVIT_CHECK_ARG(!pred || pred_stride >= n, "svit_score: pred_stride too small");
  score_kernel<<<C, kScoreBlock, 0, static_cast<cudaStream_t>(stream)>>>(logits, logits_stride, labels, n, n_cls,
                                                                         reinterpret_cast<int64_t*>(correct), loss_sum, 1,
                                                                         pred, pred_stride, accumulate);
  SVIT_LAUNCH_CHECK("score_kernel");
  return SVIT_OK;
}

extern "C" int svit_score_records(const float* logits, int64_t logits_stride, const int64_t* labels, int C, int64_t n,
                                  int n_cls, svit_record* records, int accumulate, svit_stream_t stream) {
  using namespace svit;
  SVIT_CHECK_ARG(logits && labels && records, "svit_score_records: null pointer");
  SVIT_CHECK_ARG(C >= 1 && n >= 0 && n_cls >= 1, "svit_score_records: C=%d n=%lld n_cls=%d out of range", C, (long long)n, n_cls);
  SVIT_CHECK_ARG(logits_stride >= n * n_cls, "svit_score_records: logits_stride too small");
  if ((uintptr_t)records & 15) SVIT_FAIL(SVIT_ERR_ALIGN, "svit_score_records: records must be 16-byte aligned");
  static_assert(sizeof(svit_record) == 16, "record layout");
  score_kernel<<<C, kScoreBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, logits_stride, labels, n, n_cls, &records->correct, &records->loss_sum, 2, nullptr, 0, accumulate);
  SVIT_LAUNCH_CHECK("score_kernel");
  return SVIT_OK;
}
