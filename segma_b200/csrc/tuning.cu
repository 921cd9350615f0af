// Threshold grid search on the device.
//
// scripts/tune.py:213-256 scores, for every threshold t of a grid and every label, F1 of
// (sigmoid(logits) > t) against the reference labels with sklearn, one full pass over the logits per
// threshold.  Here one pass bins every (frame, label) logit by the number of grid thresholds it exceeds
// (logit-domain cuts, bit-exact with the fp32 sigmoid rule) into two histograms per label -- reference
// positive / negative -- from which TP/FP/FN of every threshold follow by a suffix sum on the host.
// HBM-bound: 4*C + C bytes per frame.
#include <algorithm>

#include "common.cuh"

namespace segma {

constexpr int kMaxCuts = 128;

struct TuneParams {
  float cuts[kMaxCuts];
  int K;
  int C;
};

__global__ void __launch_bounds__(256) threshold_histogram_kernel(const float* __restrict__ logits,
                                                                   const uint8_t* __restrict__ truth,
                                                                   long long n_frames, TuneParams p,
                                                                   unsigned long long* __restrict__ hist) {
  extern __shared__ unsigned int s_hist[];  // [C][2][K+1]
  const int bins = p.K + 1;
  const int total = p.C * 2 * bins;
  for (int i = threadIdx.x; i < total; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  const long long n = n_frames * p.C;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = static_cast<int>(idx % p.C);
    const float x = __ldg(logits + idx);
    // number of cuts below x (cuts ascending): binary search for the first cut >= x
    int lo = 0, hi = p.K;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (x > p.cuts[mid]) lo = mid + 1; else hi = mid;
    }
    const int pos = __ldg(truth + idx) ? 1 : 0;
    atomicAdd(&s_hist[(c * 2 + pos) * bins + lo], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < total; i += blockDim.x)
    if (s_hist[i]) atomicAdd(hist + i, static_cast<unsigned long long>(s_hist[i]));
}

}  // namespace segma

using namespace segma;

extern "C" int segma_threshold_histogram(const float* logits, const uint8_t* truth, int64_t n_frames, int n_labels,
                                         const float* cuts, int n_cuts, unsigned long long* hist, void* stream) {
  SEGMA_REQUIRE(n_frames >= 0 && n_labels >= 1 && n_labels <= SEGMA_MAX_LABELS, "segma_threshold_histogram: bad shape");
  SEGMA_REQUIRE(n_cuts >= 1 && n_cuts <= kMaxCuts && cuts, "segma_threshold_histogram: 1..%d cuts", kMaxCuts);
  SEGMA_REQUIRE(hist, "segma_threshold_histogram: NULL histogram");
  for (int k = 1; k < n_cuts; ++k)
    SEGMA_REQUIRE(!(cuts[k] < cuts[k - 1]), "segma_threshold_histogram: cuts must be ascending");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t bytes = sizeof(unsigned long long) * (size_t)n_labels * 2 * (n_cuts + 1);
  SEGMA_CUDA_OK(cudaMemsetAsync(hist, 0, bytes, st));
  if (n_frames == 0) return SEGMA_OK;
  SEGMA_REQUIRE(logits && truth, "segma_threshold_histogram: NULL input");
  TuneParams p;
  p.K = n_cuts;
  p.C = n_labels;
  for (int k = 0; k < kMaxCuts; ++k) p.cuts[k] = k < n_cuts ? cuts[k] : 0.f;
  const size_t smem = sizeof(unsigned int) * (size_t)n_labels * 2 * (n_cuts + 1);
  const long long n = (long long)n_frames * n_labels;
  const int grid = (int)std::min<long long>(ceil_div_ll(n, 256 * 8), (long long)device_sm_count() * 8);
  threshold_histogram_kernel<<<std::max(grid, 1), 256, smem, st>>>(logits, truth, n_frames, p, hist);
  return launch_status("threshold_histogram_kernel");
}
