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
constexpr int kL0Threads = 128;
constexpr int kL0Row = 12;  // floats staged per time step: x[5 t .. 5 t + 9] and two of padding (48-byte rows: aligned vector loads)

// out[b][t][c] = gelu(conv(x)[t][c] * scale + shift) as fp16, time-major with `out_rows` rows per window.
// The kernel is issue bound (ncu: 72 % of the issue slots, ALU 38 % / XU 37 % / FMA 30 %), so what counts is instructions
// per output: a thread owns FOUR channels (two packed fp32 pairs: the taps are FFMA2 with the sample as a broadcast scalar
// operand, the affine and the GELU run on pairs), every time step's ten samples arrive as two 128-bit and one 64-bit
// broadcast load for all four, the four fp16 results leave as one 64-bit store (a warp writes 256 contiguous bytes), and
// the time loop advances two pointers -- 16.5 instructions per output against 31 with one pair per thread and per-step
// index arithmetic.
#ifndef SEGMA_L0_MINB
#define SEGMA_L0_MINB 6  // 80 registers: 24 warps per SM (A/B: 6.51 us per window at 100 registers, 6.36 at 80 or 72)
#endif
__global__ void __launch_bounds__(kL0Threads, SEGMA_L0_MINB) w2v2_l0_apply_kernel(
    const float* __restrict__ pcm, long long pcm_len, int win_len, long long step,
    const long long* __restrict__ win_offsets, const float* __restrict__ w,
    const float2* __restrict__ scale_shift, int C, __half* __restrict__ out, int out_rows) {
  __shared__ __align__(16) float s_x[kL0TimeTile * kL0Row];
  const int win = blockIdx.y;
  const int t0 = blockIdx.x * kL0TimeTile;
  const long long w_off = win_offsets ? win_offsets[win] : (long long)win * step;
  long long avail = pcm_len - w_off;
  if (avail > win_len) avail = win_len;
  const int T0 = avail >= kL0K ? (int)((avail - kL0K) / kL0S + 1) : 0;
  const float* x = pcm + w_off;
  for (int i = threadIdx.x; i < kL0TimeTile * kL0Row; i += kL0Threads) {
    const int tt = i / kL0Row, j = i - tt * kL0Row;
    const long long n = (long long)(t0 + tt) * kL0S + j;
    s_x[i] = (j < kL0K && n < avail) ? __ldg(x + n) : 0.f;
  }
  __syncthreads();
  const int n_rows = min(kL0TimeTile, out_rows - t0);           // rows of this tile that exist
  const int n_live = max(0, min(kL0TimeTile, T0 - t0));         // of which conv outputs (the rest is zero padding)
  for (int c = 4 * threadIdx.x; c < C; c += 4 * kL0Threads) {
    uint64_t wa[kL0K], wb[kL0K];  // taps of the channel pairs (c, c + 1) and (c + 2, c + 3)
#pragma unroll
    for (int j = 0; j < kL0K; ++j) {
      wa[j] = f2_pack(__ldg(w + c * kL0K + j), __ldg(w + (c + 1) * kL0K + j));
      wb[j] = f2_pack(__ldg(w + (c + 2) * kL0K + j), __ldg(w + (c + 3) * kL0K + j));
    }
    const float2* ssp = scale_shift + (long long)win * C + c;
    const float2 s0 = ssp[0], s1 = ssp[1], s2 = ssp[2], s3 = ssp[3];
    const uint64_t sca = f2_pack(s0.x, s1.x), sha = f2_pack(s0.y, s1.y);
    const uint64_t scb = f2_pack(s2.x, s3.x), shb = f2_pack(s2.y, s3.y);
    __half* o = out + ((long long)win * out_rows + t0) * C + c;
    const float* xs = s_x;
#pragma unroll 2
    for (int tt = 0; tt < n_live; ++tt) {
      const float4 x0 = *reinterpret_cast<const float4*>(xs);
      const float4 x1 = *reinterpret_cast<const float4*>(xs + 4);
      const float2 x2 = *reinterpret_cast<const float2*>(xs + 8);
      const float xv[kL0K] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w, x2.x, x2.y};
      uint64_t ya = f2_pack(0.f, 0.f), yb = f2_pack(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < kL0K; ++j) {
        ya = f2_fma(wa[j], f2_splat(xv[j]), ya);
        yb = f2_fma(wb[j], f2_splat(xv[j]), yb);
      }
      float a0, a1, b0, b1;
      f2_unpack(gelu_erf_pair(f2_fma(ya, sca, sha)), a0, a1);
      f2_unpack(gelu_erf_pair(f2_fma(yb, scb, shb)), b0, b1);
      SEGMA_DEV_ASSERT(c + 3 < C && t0 + tt < out_rows && t0 + tt < T0);
      *reinterpret_cast<uint2*>(o) = make_uint2(pack_f16x2(a0, a1), pack_f16x2(b0, b1));
      o += C;
      xs += kL0Row;
    }
    for (int tt = n_live; tt < n_rows; ++tt) {  // rows >= T0 are zero padding
      *reinterpret_cast<uint2*>(o) = make_uint2(0u, 0u);
      o += C;
    }
  }
}

// gate[(b*H + h)*T + i] = ga * (gb * const_h - 1) + 2, (ga, gb) = sigmoid(sum4(Linear(64 -> 8)(x[b, i, head h]))).
// The two 4-way sums commute with the projection: (ga, gb) = sigmoid(x . wa + ba, x . wb + bb) with wa / wb the sums of
// projection rows 0..3 / 4..7 -- two dot products per head instead of eight.  One warp per row: a lane reads the row's
// float4 l, l + 32, ... (coalesced; all loads of a row are in flight together), so it always sees columns 4 (l % 16) ..
// + 3 of heads 2 j + l / 16 and keeps those four (wa, wb) pairs in registers; each partial (a, b) pair is then summed over
// the 16 lanes of its half-warp with four xor shuffles.  (The first form -- eight outputs per head, every lane walking
// whole heads -- read 1.2 TB/s, stalled on dependent loads: ncu long-scoreboard 13.8 warp-cycles per issue.)
constexpr int kGateMaxHeads = 16;
constexpr int kGateRowsPerWarp = 4;  // consecutive rows per warp: the 32 weight loads of a lane are spread over them
__global__ void __launch_bounds__(256) wavlm_gate_kernel(const float* __restrict__ x, long long rows, int T, int H,
                                                          const float* __restrict__ gw, const float* __restrict__ gb,
                                                          const float* __restrict__ gconst, float* __restrict__ gate) {
  const int lane = lane_id();
  const int kc = lane & 15, half = lane >> 4;
  uint64_t wab[4];  // (wa, wb)[4 kc + i]
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float wa = 0.f, wb = 0.f;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      wa += __ldg(gw + o * 64 + 4 * kc + i);
      wb += __ldg(gw + (o + 4) * 64 + 4 * kc + i);
    }
    wab[i] = f2_pack(wa, wb);
  }
  const float ba = (__ldg(gb) + __ldg(gb + 1)) + (__ldg(gb + 2) + __ldg(gb + 3));
  const float bb = (__ldg(gb + 4) + __ldg(gb + 5)) + (__ldg(gb + 6) + __ldg(gb + 7));
  constexpr int kMaxJ = kGateMaxHeads / 2;
  const int h = 2 * kc + half;  // lane kc of a half-warp writes head 2 kc + half
  const float gc = (kc < kMaxJ && h < H) ? __ldg(gconst + h) : 0.f;
  const long long row0 = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * kGateRowsPerWarp;
#pragma unroll 1
  for (long long row = row0; row < min(rows, row0 + kGateRowsPerWarp); ++row) {
    const long long b = row / T;
    const int i = (int)(row - b * T);
    const float4* xr = reinterpret_cast<const float4*>(x + row * (long long)(H * 64));
    float4 a[kMaxJ];
#pragma unroll
    for (int j = 0; j < kMaxJ; ++j)
      a[j] = (2 * j + half < H) ? __ldg(xr + 32 * j + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    float mine_a = 0.f, mine_b = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxJ; ++j) {
      if (2 * j >= H) break;  // warp-uniform
      uint64_t acc = f2_fma(wab[0], f2_splat(a[j].x), f2_pack(0.f, 0.f));
      acc = f2_fma(wab[1], f2_splat(a[j].y), acc);
      acc = f2_fma(wab[2], f2_splat(a[j].z), acc);
      acc = f2_fma(wab[3], f2_splat(a[j].w), acc);
      float va, vb;
      f2_unpack(acc, va, vb);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {  // over the 16 lanes of the half-warp (head 2 j + half)
        va += __shfl_xor_sync(0xffffffffu, va, o);
        vb += __shfl_xor_sync(0xffffffffu, vb, o);
      }
      if (kc == j) { mine_a = va; mine_b = vb; }
    }
    if (kc < kMaxJ && h < H) {
      const float ga = 1.0f / (1.0f + expf(-(mine_a + ba))), gbv = 1.0f / (1.0f + expf(-(mine_b + bb)));
      SEGMA_DEV_ASSERT(row < rows && h < H && i < T);
      gate[(b * H + h) * T + i] = ga * (gbv * gc - 1.0f) + 2.0f;
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
  SEGMA_REQUIRE(channels > 0 && channels % 4 == 0, "segma_w2v2_layer0: channels must be a multiple of 4");
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
  SEGMA_REQUIRE(n_heads <= kGateMaxHeads, "segma_wavlm_gate: at most %d heads", kGateMaxHeads);
  wavlm_gate_kernel<<<(unsigned)ceil_div_ll(rows, 8 * kGateRowsPerWarp), 256, 0, (cudaStream_t)stream>>>(x, rows, T, n_heads, gate_w,
                                                                                      gate_b, gate_const, gate);
  return launch_status("wavlm_gate_kernel");
}

}  // extern "C"
