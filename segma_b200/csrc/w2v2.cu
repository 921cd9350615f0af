// wav2vec2 / HuBERT / WavLM waveform front end, layer 0, and the WavLM relative-position gate.
//
// Replaces the first ConvLayerBlock of torchaudio's FeatureExtractor -- Conv1d(1, C, k=10, s=5, no bias) ->
// GroupNorm(C groups, i.e. per (window, channel) over time) -> exact GELU
// (site-packages/torchaudio/models/wav2vec2/components.py:77-99,117-143) -- as called from
// SurgicalHydraHubert.forward (src/segma/models/hubert/surgical_hydra.py:88-89), and the gate of
// WavLMSelfAttention (site-packages/torchaudio/models/wav2vec2/wavlm_attention.py:185-193).
//
// The layer-0 output (C x 12 799 per 4 s window, 26 MB in fp32) is never stored un-normalised: the
// GroupNorm statistics follow from second moments of the *input* (y_c = w_c . x[5t : 5t+10] is linear, so
// mean_c = w_c . m and var_c = w_c^T Cov w_c with the 10-vector m and the 10x10 covariance of the strided
// input patches), which one pass over the 256 KB window yields in fp64; a second kernel then writes
// gelu(GroupNorm(conv)) directly as the fp16 time-major activation the layer-1 implicit GEMM reads.
#include "common.cuh"

namespace segma {

constexpr int kL0K = 10, kL0S = 5;
constexpr int kStatsThreads = 256;
constexpr int kPairs = kL0K * (kL0K + 1) / 2;  // 55

__global__ void __launch_bounds__(kStatsThreads) w2v2_l0_stats_kernel(
    const float* __restrict__ pcm, long long pcm_len, int win_len, long long step,
    const long long* __restrict__ win_offsets, const float* __restrict__ w,
    const float* __restrict__ gamma, const float* __restrict__ beta, int C, float2* __restrict__ scale_shift) {
  __shared__ double s_sum[kL0K + kPairs];
  __shared__ double s_part[kStatsThreads / 32][kL0K + kPairs];
  const int win = blockIdx.x;
  // window w starts at sample w * step, or at win_offsets[w] when windows of several files are packed into one call
  const long long w_off = win_offsets ? win_offsets[win] : (long long)win * step;
  long long avail = pcm_len - w_off;
  if (avail > win_len) avail = win_len;
  const int T0 = avail >= kL0K ? (int)((avail - kL0K) / kL0S + 1) : 0;
  const float* x = pcm + w_off;
  double acc[kL0K + kPairs];
#pragma unroll
  for (int i = 0; i < kL0K + kPairs; ++i) acc[i] = 0.0;
  // fp32 partial sums over short runs, folded into fp64 every 32 patches
  for (int t0 = threadIdx.x * 32; t0 < T0; t0 += kStatsThreads * 32) {
    float part[kL0K + kPairs];
#pragma unroll
    for (int i = 0; i < kL0K + kPairs; ++i) part[i] = 0.f;
    const int t1 = min(t0 + 32, T0);
    for (int t = t0; t < t1; ++t) {
      float v[kL0K];
#pragma unroll
      for (int j = 0; j < kL0K; ++j) v[j] = __ldg(x + (long long)t * kL0S + j);
      int q = kL0K;
#pragma unroll
      for (int j = 0; j < kL0K; ++j) {
        part[j] += v[j];
#pragma unroll
        for (int k = j; k < kL0K; ++k) {
          part[q] = fmaf(v[j], v[k], part[q]);
          ++q;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kL0K + kPairs; ++i) acc[i] += (double)part[i];
  }
#pragma unroll
  for (int i = 0; i < kL0K + kPairs; ++i) {
    double v = acc[i];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane_id() == 0) s_part[threadIdx.x >> 5][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < kL0K + kPairs) {
    double v = 0.0;
    for (int wi = 0; wi < kStatsThreads / 32; ++wi) v += s_part[wi][threadIdx.x];
    s_sum[threadIdx.x] = v;
  }
  __syncthreads();
  const double inv_n = T0 > 0 ? 1.0 / T0 : 0.0;
  for (int c = threadIdx.x; c < C; c += kStatsThreads) {
    double wc[kL0K];
#pragma unroll
    for (int j = 0; j < kL0K; ++j) wc[j] = (double)__ldg(w + c * kL0K + j);
    double mean = 0.0;
#pragma unroll
    for (int j = 0; j < kL0K; ++j) mean += wc[j] * s_sum[j] * inv_n;
    double ey2 = 0.0;
    int q = kL0K;
#pragma unroll
    for (int j = 0; j < kL0K; ++j)
#pragma unroll
      for (int k = j; k < kL0K; ++k) {
        const double cov = s_sum[q++] * inv_n - (s_sum[j] * inv_n) * (s_sum[k] * inv_n);
        ey2 += (j == k ? 1.0 : 2.0) * wc[j] * wc[k] * cov;
      }
    const double var = ey2 > 0.0 ? ey2 : 0.0;  // biased variance of the conv output
    const double rstd = 1.0 / sqrt(var + 1e-5);
    const double g = (double)__ldg(gamma + c), bta = (double)__ldg(beta + c);
    scale_shift[(long long)win * C + c] = make_float2((float)(g * rstd), (float)(bta - mean * g * rstd));
  }
}

constexpr int kL0TimeTile = 64;
constexpr int kL0Threads = 256;

// out[b][t][c] = gelu(conv(x)[t][c] * scale + shift) as fp16, time-major with `out_rows` rows per window
__global__ void __launch_bounds__(kL0Threads) w2v2_l0_apply_kernel(
    const float* __restrict__ pcm, long long pcm_len, int win_len, long long step,
    const long long* __restrict__ win_offsets, const float* __restrict__ w,
    const float2* __restrict__ scale_shift, int C, __half* __restrict__ out, int out_rows) {
  __shared__ float s_x[kL0TimeTile * kL0S + kL0K];
  const int win = blockIdx.y;
  const int t0 = blockIdx.x * kL0TimeTile;
  const long long w_off = win_offsets ? win_offsets[win] : (long long)win * step;
  long long avail = pcm_len - w_off;
  if (avail > win_len) avail = win_len;
  const int T0 = avail >= kL0K ? (int)((avail - kL0K) / kL0S + 1) : 0;
  const float* x = pcm + w_off;
  for (int i = threadIdx.x; i < kL0TimeTile * kL0S + kL0K; i += kL0Threads) {
    const long long n = (long long)t0 * kL0S + i;
    s_x[i] = n < avail ? __ldg(x + n) : 0.f;
  }
  __syncthreads();
  __half* o = out + (long long)win * out_rows * C;
  for (int c2 = threadIdx.x; c2 < C / 2; c2 += kL0Threads) {
    const int c = c2 * 2;
    // the two channels ride in one packed fp32 pair: FFMA2 for the taps and the affine, gelu_erf_pair for the GELU
    uint64_t w2[kL0K];
#pragma unroll
    for (int j = 0; j < kL0K; ++j) w2[j] = f2_pack(__ldg(w + c * kL0K + j), __ldg(w + (c + 1) * kL0K + j));
    const float2 ssa = scale_shift[(long long)win * C + c], ssb = scale_shift[(long long)win * C + c + 1];
    const uint64_t sc2 = f2_pack(ssa.x, ssb.x), sh2 = f2_pack(ssa.y, ssb.y);
    for (int tt = 0; tt < kL0TimeTile; ++tt) {
      const int t = t0 + tt;
      if (t >= out_rows) break;
      float ya = 0.f, yb = 0.f;
      if (t < T0) {
        uint64_t y2 = f2_pack(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < kL0K; ++j) y2 = f2_fma(w2[j], f2_splat(s_x[tt * kL0S + j]), y2);
        f2_unpack(gelu_erf_pair(f2_fma(y2, sc2, sh2)), ya, yb);
      }
      *reinterpret_cast<uint32_t*>(o + (long long)t * C + c) = pack_f16x2(ya, yb);  // rows >= T0 are zero padding
    }
  }
}

// gate[(b*H + h)*T + i] = ga * (gb * const_h - 1) + 2, (ga, gb) = sigmoid(sum4(Linear(64 -> 8)(x[b, i, head h])))
// One warp per row; lane = (head mod 4, output): its 64-long dot products for heads h, h + 4, h + 8, ... share one
// projection row, read from shared memory once per 4 columns and reused for every head (rows padded to 68 floats:
// the 8 lanes of a quarter warp read 8 different rows at the same column, which a 64-float pitch would put in the
// same four banks).  The 8 lanes of a group read the same head slice of x (one broadcast transaction).  The two
// 4-way sums are two shuffles inside aligned groups of 8 lanes.
constexpr int kGateMaxPasses = 4;  // up to 16 heads
__global__ void __launch_bounds__(256) wavlm_gate_kernel(const float* __restrict__ x, long long rows, int T, int H,
                                                          const float* __restrict__ gw, const float* __restrict__ gb,
                                                          const float* __restrict__ gconst, float* __restrict__ gate) {
  constexpr int kPitch = 68;
  __shared__ __align__(16) float s_w[8 * kPitch];
  __shared__ float s_b[8];
  for (int i = threadIdx.x; i < 8 * 64; i += blockDim.x) s_w[(i >> 6) * kPitch + (i & 63)] = gw[i];
  if (threadIdx.x < 8) s_b[threadIdx.x] = gb[threadIdx.x];
  __syncthreads();
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = lane_id();
  const int o = lane & 7, h0 = lane >> 3;
  const long long b = row / T;
  const int i = (int)(row - b * T);
  const float4* xr = reinterpret_cast<const float4*>(x + row * (long long)(H * 64));
  const float4* wv = reinterpret_cast<const float4*>(s_w + o * kPitch);
  float acc[kGateMaxPasses];
#pragma unroll
  for (int p = 0; p < kGateMaxPasses; ++p) acc[p] = 0.f;
#pragma unroll 4
  for (int k = 0; k < 16; ++k) {
    const float4 w4 = wv[k];
#pragma unroll
    for (int p = 0; p < kGateMaxPasses; ++p) {
      const int h = h0 + 4 * p;
      if (h < H) {
        const float4 a = __ldg(xr + h * 16 + k);
        acc[p] = fmaf(a.x, w4.x, acc[p]);
        acc[p] = fmaf(a.y, w4.y, acc[p]);
        acc[p] = fmaf(a.z, w4.z, acc[p]);
        acc[p] = fmaf(a.w, w4.w, acc[p]);
      }
    }
  }
#pragma unroll
  for (int p = 0; p < kGateMaxPasses; ++p) {
    const int h = h0 + 4 * p;
    if (4 * p >= H) break;  // warp-uniform
    float v = acc[p] + s_b[o];
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);                  // lanes o = 0 and o = 4 hold the two 4-way sums
    const float other = __shfl_down_sync(0xffffffffu, v, 4);
    if (h < H && o == 0) {
      const float ga = 1.0f / (1.0f + expf(-v)), gbv = 1.0f / (1.0f + expf(-other));
      gate[(b * H + h) * T + i] = ga * (gbv * __ldg(gconst + h) - 1.0f) + 2.0f;
    }
  }
}

}  // namespace segma

using namespace segma;

extern "C" {

static int w2v2_layer0_impl(const float* pcm, int64_t pcm_len, int n_windows, int win_len, int64_t step,
                           const int64_t* win_offsets, const float* w, const float* gamma, const float* beta, int channels,
                           void* scale_shift, void* out, int out_rows, void* stream) {
  SEGMA_REQUIRE(n_windows >= 0, "segma_w2v2_layer0: negative n_windows");
  if (n_windows == 0) return SEGMA_OK;
  SEGMA_REQUIRE(pcm && w && gamma && beta && scale_shift && out, "segma_w2v2_layer0: NULL buffer");
  SEGMA_REQUIRE(channels > 0 && channels % 2 == 0, "segma_w2v2_layer0: channels must be even");
  SEGMA_REQUIRE(win_len >= kL0K && step >= 0 && pcm_len >= 0, "segma_w2v2_layer0: bad window geometry");
  const int T0 = (win_len - kL0K) / kL0S + 1;
  SEGMA_REQUIRE(out_rows >= T0, "segma_w2v2_layer0: out_rows %d < %d conv outputs", out_rows, T0);
  SEGMA_REQUIRE(n_windows <= 65535, "segma_w2v2_layer0: at most 65535 windows per call");
  cudaStream_t st = (cudaStream_t)stream;
  const long long* offs = reinterpret_cast<const long long*>(win_offsets);
  w2v2_l0_stats_kernel<<<n_windows, kStatsThreads, 0, st>>>(pcm, pcm_len, win_len, step, offs, w, gamma, beta, channels,
                                                            static_cast<float2*>(scale_shift));
  int rc = launch_status("w2v2_l0_stats_kernel");
  if (rc != SEGMA_OK) return rc;
  dim3 grid(ceil_div(out_rows, kL0TimeTile), n_windows);
  w2v2_l0_apply_kernel<<<grid, kL0Threads, 0, st>>>(pcm, pcm_len, win_len, step, offs, w,
                                                    static_cast<const float2*>(scale_shift), channels,
                                                    static_cast<__half*>(out), out_rows);
  return launch_status("w2v2_l0_apply_kernel");
}

int segma_w2v2_layer0(const float* pcm, int64_t pcm_len, int n_windows, int win_len, int64_t step, const float* w,
                      const float* gamma, const float* beta, int channels, void* scale_shift, void* out,
                      int out_rows, void* stream) {
  return w2v2_layer0_impl(pcm, pcm_len, n_windows, win_len, step, nullptr, w, gamma, beta, channels, scale_shift, out,
                          out_rows, stream);
}

int segma_w2v2_layer0_at(const float* pcm, int64_t pcm_len, int n_windows, int win_len, const int64_t* win_offsets,
                         const float* w, const float* gamma, const float* beta, int channels, void* scale_shift,
                         void* out, int out_rows, void* stream) {
  SEGMA_REQUIRE(win_offsets != nullptr, "segma_w2v2_layer0_at: NULL win_offsets");
  return w2v2_layer0_impl(pcm, pcm_len, n_windows, win_len, 0, win_offsets, w, gamma, beta, channels, scale_shift, out,
                          out_rows, stream);
}

int segma_wavlm_gate(const float* x, int64_t rows, int T, int n_heads, const float* gate_w, const float* gate_b,
                     const float* gate_const, float* gate, void* stream) {
  SEGMA_REQUIRE(rows >= 0 && T > 0 && n_heads > 0 && rows % T == 0, "segma_wavlm_gate: bad shape");
  if (rows == 0) return SEGMA_OK;
  SEGMA_REQUIRE(x && gate_w && gate_b && gate_const && gate, "segma_wavlm_gate: NULL buffer");
  SEGMA_REQUIRE(n_heads <= 4 * kGateMaxPasses, "segma_wavlm_gate: at most %d heads", 4 * kGateMaxPasses);
  wavlm_gate_kernel<<<(unsigned)ceil_div_ll(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, rows, T, n_heads, gate_w,
                                                                                      gate_b, gate_const, gate);
  return launch_status("wavlm_gate_kernel");
}

}  // extern "C"
