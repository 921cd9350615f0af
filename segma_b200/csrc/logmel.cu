// Fused windowing + Whisper log-mel front end (fp32).
//
// Replaces `unfold` (src/segma/inference.py:148-152) + WhisperFeatureExtractor
// (site-packages/transformers/models/whisper/feature_extraction_whisper.py:135-164): for every 4 s window,
// STFT frames of 400 samples at hop 160 (periodic Hann, centre + reflect padding of the 30 s padded window)
// -> 201-bin power spectrum -> 80 slaney mel bins (391 non-zeros) -> log10(max(., 1e-10))
// -> max(., window_max - 8) -> (. + 4) / 4, padded with the constant value to 3000 frames.
//
// Per window 256 KB of PCM in, 960 KB of features out (831 KB of it the constant tail).  One kernel: a cluster of four
// CTAs per window stages the overlapping samples of 16 frames at a time in shared memory (cp.async, one group ahead),
// transforms two frames per 400-point complex FFT done as 20 x 20 in registers, applies the sparse mel filters and the
// log, and keeps the unclamped values in an L2-resident scratch; the cluster's maxima meet through distributed shared
// memory and each CTA then writes a quarter of the window -- clamp, scale, constant tail, optionally also the fp16
// time-major tile the conv-stem GEMM consumes -- while the cluster already transforms its next window.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace segma {

constexpr int kNfft = 400;
constexpr int kHop = 160;
constexpr int kBins = 201;
constexpr int kMels = SEGMA_MEL_BINS;
constexpr int kFramesOut = SEGMA_MEL_FRAMES;
constexpr int kPadSamples = 480000;  // 30 s
constexpr int kGroup = 16;           // frames per work item: 8 frame pairs x 20 threads
constexpr int kMaxTaps = 32;         // max contiguous FFT bins per mel filter
constexpr int kMelTapsSmem = 15;     // taps per filter kept in shared memory (the slaney bank has at most 14; wider banks read global memory)
constexpr int kStage = (kGroup - 1) * kHop + kNfft;  // 2800 staged samples per group

struct MelTables {
  float hann[kNfft];
  int mel_k0[kMels];
  int mel_len[kMels];
  float mel_w[kMels][kMaxTaps];
  float2 tw_b1[20], tw_b4[20];     // exp(-2 pi i b / 400) and its fourth power: the twiddles W400^{b c} between the
                                   // two radix-20 passes are products of these
  float2 mel_w2[kMels][kMaxTaps];  // (w / 4, w / 4): filter taps for a pair of frames (the 1/4 of the unpacked power)
};

__device__ MelTables g_tab;

static std::mutex g_tab_mutex;
static PerDeviceFlag g_tab_ready;  // the __device__ tables exist once per device
static std::vector<float> g_mel_dense;  // (201, 80) row-major, active filterbank

static double hz_to_mel(double f) {
  return f >= 1000.0 ? 15.0 + std::log(f / 1000.0) * (27.0 / std::log(6.4)) : 3.0 * f / 200.0;
}
static double mel_to_hz(double m) {
  return m >= 15.0 ? 1000.0 * std::exp((std::log(6.4) / 27.0) * (m - 15.0)) : 200.0 * m / 3.0;
}

// slaney-scale, slaney-normalised triangular filters, 0-8 kHz, 16 kHz sampling (float64, then fp32)
static void default_mel(std::vector<float>& dense) {
  dense.assign((size_t)kBins * kMels, 0.f);
  double edges[kMels + 2];
  const double m_lo = hz_to_mel(0.0), m_hi = hz_to_mel(8000.0);
  for (int i = 0; i < kMels + 2; ++i) edges[i] = mel_to_hz(m_lo + (m_hi - m_lo) * i / (kMels + 1));
  for (int k = 0; k < kBins; ++k) {
    const double f = 8000.0 * k / (kBins - 1);
    for (int m = 0; m < kMels; ++m) {
      const double down = (f - edges[m]) / (edges[m + 1] - edges[m]);
      const double up = (edges[m + 2] - f) / (edges[m + 2] - edges[m + 1]);
      const double v = std::max(0.0, std::min(down, up)) * (2.0 / (edges[m + 2] - edges[m]));
      dense[(size_t)k * kMels + m] = (float)v;
    }
  }
}

static int upload_tables_locked() {
  static MelTables h;  // large: keep off the stack
  std::memset(&h, 0, sizeof(h));
  const double two_pi = 6.283185307179586476925286766559;
  for (int n = 0; n < kNfft; ++n) h.hann[n] = (float)(0.5 - 0.5 * std::cos(two_pi * n / kNfft));
  for (int m = 0; m < kMels; ++m) {
    int first = -1, last = -1;
    for (int k = 0; k < kBins; ++k)
      if (g_mel_dense[(size_t)k * kMels + m] != 0.f) {
        if (first < 0) first = k;
        last = k;
      }
    if (first < 0) { first = 0; last = -1; }
    const int len = last - first + 1;
    if (len > kMaxTaps) {
      set_last_error("mel filter %d spans %d FFT bins (max %d)", m, len, kMaxTaps);
      return SEGMA_ERR_UNSUPPORTED;
    }
    h.mel_k0[m] = first;
    h.mel_len[m] = len;
    for (int i = 0; i < len; ++i) {
      h.mel_w[m][i] = g_mel_dense[(size_t)(first + i) * kMels + m];
      h.mel_w2[m][i] = make_float2(0.25f * h.mel_w[m][i], 0.25f * h.mel_w[m][i]);
    }
  }
  // Step 4 of the kernel: the 8 filters x 2 pair groups of a half-warp read (P_A, P_B)[k0 + i] in lockstep and collide on
  // an 8-byte bank iff two first bins agree mod 8 (the log-spaced filters do, often).  A filter may start up to a few
  // bins early with zero taps in front -- free while it stays within its warp's longest filter (+ 1) -- which moves
  // its bank: longest filters first, each takes the smallest shift that gives it an unused residue.
  for (int w0 = 0; w0 < kMels; w0 += 16) {
    int longest = 0;
    for (int m = w0; m < std::min(kMels, w0 + 16); ++m) longest = std::max(longest, h.mel_len[m]);
    for (int g0 = w0; g0 < std::min(kMels, w0 + 16); g0 += 8) {
      int ms[8], n = 0;
      for (int m = g0; m < std::min(kMels, g0 + 8); ++m) ms[n++] = m;
      std::stable_sort(ms, ms + n, [&](int a, int b2) { return h.mel_len[a] > h.mel_len[b2]; });
      bool used[8] = {};
      for (int i = 0; i < n; ++i) {
        const int m = ms[i];
        const int smax = std::min({h.mel_k0[m], longest + 1 - h.mel_len[m], kMelTapsSmem - h.mel_len[m]});
        int shift = 0;
        for (int sft = 0; sft <= smax; ++sft)
          if (!used[(h.mel_k0[m] - sft) & 7]) { shift = sft; break; }
        if (shift > 0) {
          for (int t = h.mel_len[m] - 1; t >= 0; --t) {
            h.mel_w[m][t + shift] = h.mel_w[m][t];
            h.mel_w2[m][t + shift] = h.mel_w2[m][t];
          }
          for (int t = 0; t < shift; ++t) {
            h.mel_w[m][t] = 0.f;
            h.mel_w2[m][t] = make_float2(0.f, 0.f);
          }
          h.mel_k0[m] -= shift;
          h.mel_len[m] += shift;
        }
        used[h.mel_k0[m] & 7] = true;
      }
    }
  }
  for (int b = 0; b < 20; ++b) {
    h.tw_b1[b] = make_float2((float)std::cos(two_pi * b / kNfft), (float)-std::sin(two_pi * b / kNfft));
    h.tw_b4[b] = make_float2((float)std::cos(two_pi * 4 * b / kNfft), (float)-std::sin(two_pi * 4 * b / kNfft));
  }
  SEGMA_CUDA_OK(cudaMemcpyToSymbol(g_tab, &h, sizeof(h)));
  g_tab_ready.here() = true;
  return SEGMA_OK;
}

static int ensure_tables() {
  std::lock_guard<std::mutex> lock(g_tab_mutex);
  if (g_tab_ready.here()) return SEGMA_OK;
  if (g_mel_dense.empty()) default_mel(g_mel_dense);
  return upload_tables_locked();
}

// ---- device helpers --------------------------------------------------------------------------------
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
// padded-window sample n (after torch.stft's reflect padding of the 480000-sample buffer)
__device__ __forceinline__ float sample_at(const float* __restrict__ w, long long n, long long avail) {
  if (n < 0) n = -n;
  if (n >= kPadSamples) n = 2ll * (kPadSamples - 1) - n;
  return (n < avail) ? __ldg(w + n) : 0.f;
}

// ---- fused kernel: one cluster of four CTAs per window ------------------------------------------------------
// Transform.  Two frames A, B ride one 400-point complex FFT (z = h x_A + i h x_B) done as 20 x 20: with n = 20 a + b
// and k = c + 20 d,  Z[c + 20 d] = sum_b W20^{bd} ( W400^{bc} sum_a W20^{ac} z[20 a + b] ).  A thread owns one
// 20-point DFT per pass (Good-Thomas 4 x 5: no inner twiddles, all in registers), so a frame pair needs one exchange
// through shared memory between the passes and one to bring Z[k] and Z[400 - k] together for
// |X_A[k]|^2 = |Z[k] + conj Z[400-k]|^2 / 4,  |X_B[k]|^2 = |Z[k] - conj Z[400-k]|^2 / 4  (the 1/4 sits in the mel taps).
// A group of 16 frames = 8 pairs x 20 threads = 160 threads, every thread busy in both passes.
// Finish.  The four CTAs of a cluster take the window's frame groups in turn, leave the unclamped log10 mel values in
// the (L2-resident) scratch, exchange their maxima through distributed shared memory and then each write a quarter
// of the window's output -- clamp, scale and the constant tail -- so that one window's stores overlap the other
// resident clusters' transforms and no second kernel re-reads anything from HBM.
#ifndef SEGMA_LOGMEL_CL
#define SEGMA_LOGMEL_CL 4
#endif
constexpr int kCl = SEGMA_LOGMEL_CL;                         // CTAs per window
constexpr int kPairs = kGroup / 2;             // frame pairs per group
constexpr int kThreadsF = kPairs * 20;         // 160
constexpr int kChunkPad = kHop + 10;           // staged samples: 160-sample chunks 170 words apart, which puts the
                                               // 20 lanes of pair p on banks 20 p + b (conflict-free pass-1 loads)
constexpr int kStageChunks = (kStage + kHop - 1) / kHop;
constexpr int kStageWords = kStageChunks * kChunkPad;
constexpr int kEL = 21;                        // exchange rows [c][b], odd stride: pass 2 reads columns conflict-free
constexpr int kEP = 20 * kEL;                  // 420 = 4 mod 16: the 8 pairs of a warp fall on distinct 8-byte banks
constexpr int kZP = 404;                       // spectrum of a pair, natural order, same residue
constexpr int kPP = kBins + 1;                 // (P_A, P_B)[k] of a pair; 4 kPP = 8 mod 16: the two pair groups of a
                                               // mel filter fall on different banks
constexpr int kTmTileF = 32;
constexpr int kR1Bytes = kPairs * kEP * 8;     // stage / exchange / spectrum / transpose tile share one region
static_assert(kPairs * kZP * 8 <= kR1Bytes && kMels * (kTmTileF + 1) * 4 <= kR1Bytes,
              "shared region too small");
static_assert(kMels % kCl == 0 && kThreadsF % 20 == 0, "row split");
static_assert(kThreadsF == 160 && kBins == 201, "bin mapping of the power pass");
static_assert(kR1Bytes % 16 == 0 && (kStageWords * 4) % 16 == 0, "shared-memory carve-up alignment");
constexpr int kFusedSmem = kR1Bytes + kStageWords * 4 + kPairs * kPP * 8 + kMels * kMelTapsSmem * 4;

// Complex numbers as packed fp32 pairs (re, im): additions, real scalings and fused multiply-adds of a whole complex
// number are one FADD2 / FMUL2 / FFMA2 (two fp32 lanes per issue slot on sm_100); multiplications by -i are written on
// the scalar halves and cost nothing extra.
typedef uint64_t cpx;
__device__ __forceinline__ cpx cpx_make(float re, float im) { return f2_pack(re, im); }
__device__ __forceinline__ cpx cpx_sub(cpx a, cpx b) {
  cpx r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

__device__ __forceinline__ void dft4(cpx& x0, cpx& x1, cpx& x2, cpx& x3) {
  const cpx t0 = f2_add(x0, x2), t1 = cpx_sub(x0, x2), t2 = f2_add(x1, x3), t3 = cpx_sub(x1, x3);
  x0 = f2_add(t0, t2);
  x2 = cpx_sub(t0, t2);
  float t1x, t1y, t3x, t3y;
  f2_unpack(t1, t1x, t1y);
  f2_unpack(t3, t3x, t3y);
  x1 = cpx_make(t1x + t3y, t1y - t3x);   // t1 - i t3
  x3 = cpx_make(t1x - t3y, t1y + t3x);   // t1 + i t3
}

__device__ __forceinline__ void dft5(cpx (&v)[5]) {
  const float c1 = 0.30901699437494742410f;   // cos(2pi/5)
  const float c2 = -0.80901699437494742410f;  // cos(4pi/5)
  const float s1 = 0.95105651629515357212f;   // sin(2pi/5)
  const float s2 = 0.58778525229247312917f;   // sin(4pi/5)
  float x1, y1, x2, y2, x3, y3, x4, y4;
  f2_unpack(v[1], x1, y1);
  f2_unpack(v[2], x2, y2);
  f2_unpack(v[3], x3, y3);
  f2_unpack(v[4], x4, y4);
  const cpx a1 = f2_add(v[1], v[4]), a2 = f2_add(v[2], v[3]);
  const cpx r1 = cpx_make(y1 - y4, x4 - x1), r2 = cpx_make(y2 - y3, x3 - x2);  // -i (v1 - v4), -i (v2 - v3)
  const cpx x0 = v[0];
  v[0] = f2_add(f2_add(x0, a1), a2);
  const cpx p1 = f2_fma(f2_splat(c2), a2, f2_fma(f2_splat(c1), a1, x0));
  const cpx p2 = f2_fma(f2_splat(c1), a2, f2_fma(f2_splat(c2), a1, x0));
  const cpx q1 = f2_fma(f2_splat(s2), r2, f2_mul(f2_splat(s1), r1));   // -i (s1 b1 + s2 b2)
  const cpx q2 = f2_fma(f2_splat(-s1), r2, f2_mul(f2_splat(s2), r1));  // -i (s2 b1 - s1 b2)
  v[1] = f2_add(p1, q1);
  v[4] = cpx_sub(p1, q1);
  v[2] = f2_add(p2, q2);
  v[3] = cpx_sub(p2, q2);
}

// forward DFT of length 20, natural order in and out: a = 5 a1 + 4 a2, c = 5 c1 + 16 c2 (mod 20), so that
// W20^{ac} = W4^{a1 c1} W5^{a2 c2} -- four DFT-5 and five DFT-4 without twiddles; the index maps are register renaming
__device__ __forceinline__ void dft20(cpx (&v)[20]) {
  cpx u[4][5];
#pragma unroll
  for (int a1 = 0; a1 < 4; ++a1)
#pragma unroll
    for (int a2 = 0; a2 < 5; ++a2) u[a1][a2] = v[(5 * a1 + 4 * a2) % 20];
#pragma unroll
  for (int a1 = 0; a1 < 4; ++a1) dft5(u[a1]);
#pragma unroll
  for (int c2 = 0; c2 < 5; ++c2) dft4(u[0][c2], u[1][c2], u[2][c2], u[3][c2]);
#pragma unroll
  for (int c1 = 0; c1 < 4; ++c1)
#pragma unroll
    for (int c2 = 0; c2 < 5; ++c2) v[(5 * c1 + 16 * c2) % 20] = u[c1][c2];
}

__device__ __forceinline__ float2 cpx_f2(cpx a) {
  float2 r;
  f2_unpack(a, r.x, r.y);
  return r;
}

// the unclamped log-mel rows wait in L2 for the finish one window later: written with an evict-last policy, and
// dropped from L2 without a write-back once they have been read (they never need to reach HBM)
__device__ __forceinline__ uint64_t l2_keep_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void st_f32x2_keep(float* p, float a, float b, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(a), "f"(b), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_f32x8_keep(float* p, const float (&v)[8], uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8}, %9;" ::"l"(p), "f"(v[0]), "f"(v[1]),
               "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void l2_discard_128(const void* p) {
  asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
}

__device__ __forceinline__ float ld_cluster_f32(const float* own_smem, uint32_t rank) {
  uint32_t remote;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(own_smem)), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
  return v;
}

__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

// stage samples [s0, s0 + 2800) of the window into 160-sample chunks 170 words apart.  Interior groups of an 8-byte
// aligned window go through cp.async (no registers, in flight while the previous group is transformed); groups that
// touch the reflected start, the end of the audio or an odd address take the scalar path.
__device__ __forceinline__ void stage_group(float* stage, const float* __restrict__ w, long long s0, long long avail,
                                            bool aligned8, int tid) {
  if (aligned8 && s0 >= 0 && s0 + kStage <= avail) {
    const float* src = w + s0 + 2 * tid;
    float* dst = stage + (tid / 80) * kChunkPad + 2 * (tid % 80);
#pragma unroll
    for (int it = 0; it < (kStage / 2 + kThreadsF - 1) / kThreadsF; ++it) {
      if (it * kThreadsF + tid < kStage / 2) cp_async_8(dst + it * 2 * kChunkPad, src + it * 2 * kThreadsF);
    }
  } else {
    for (int q = tid; q < kStage / 2; q += kThreadsF) {
      const long long n = s0 + 2 * q;
      const int chunk = q / (kHop / 2);
      *reinterpret_cast<float2*>(stage + chunk * kChunkPad + 2 * (q - chunk * (kHop / 2))) =
          make_float2(sample_at(w, n, avail), sample_at(w, n + 1, avail));
    }
  }
}

__device__ __forceinline__ int window_valid_frames(long long pcm_len, int win, long long step, int win_len,
                                                   long long& avail) {
  avail = pcm_len - (long long)win * step;
  if (avail > win_len) avail = win_len;
  if (avail <= 0) { avail = 0; return 0; }
  const long long nv = (avail + 200 + kHop - 1) / kHop;  // frames >= nv see only zeros
  return nv > kFramesOut ? kFramesOut : (int)nv;
}

// clamp / scale / constant tail of one window: this CTA's quarter of the outputs
__device__ __forceinline__ void finish_window(int win, uint32_t rank, int tid, int n_valid, float gmax, int nvp,
                                              const float* __restrict__ ls, float* __restrict__ out_f32,
                                              __half* __restrict__ out_tm, unsigned char* r1) {
  if (gmax == -INFINITY) gmax = -10.f;                  // no frame touches audio
  if (n_valid < kFramesOut) gmax = fmaxf(gmax, -10.f);  // the zero padding is part of the reference's maximum
  const float floor_v = gmax - 8.0f;
  const float fill = (fmaxf(-10.f, floor_v) + 4.0f) / 4.0f;
  if (out_f32) {
    constexpr int kRows = kMels / kCl, kQuads = kFramesOut / 4;
    const int nq = (n_valid + 3) / 4;
    float4* o = reinterpret_cast<float4*>(out_f32 + ((long long)win * kMels + rank * kRows) * kFramesOut);
    const float4* src = reinterpret_cast<const float4*>(ls + (long long)rank * kRows * nvp);
    const int nvq = nvp / 4;
    // frames that touch audio: clamp and scale what the cluster left in the scratch (four rows in flight per thread)
    for (int q = tid; q < nq; q += kThreadsF) {
      const int t = 4 * q;
      constexpr int kInFlight = kRows % 4 == 0 ? 4 : (kRows % 5 == 0 ? 5 : 1);  // rows per thread in flight
      static_assert(kRows % kInFlight == 0, "rows per CTA");
#pragma unroll 1
      for (int row = 0; row < kRows; row += kInFlight) {
        float4 x[kInFlight];
#pragma unroll
        for (int i = 0; i < kInFlight; ++i) x[i] = __ldcg(src + (row + i) * nvq + q);
#pragma unroll
        for (int i = 0; i < kInFlight; ++i) {
          float4 r = make_float4(fill, fill, fill, fill);
          r.x = (fmaxf(x[i].x, floor_v) + 4.0f) / 4.0f;
          if (t + 1 < n_valid) r.y = (fmaxf(x[i].y, floor_v) + 4.0f) / 4.0f;
          if (t + 2 < n_valid) r.z = (fmaxf(x[i].z, floor_v) + 4.0f) / 4.0f;
          if (t + 3 < n_valid) r.w = (fmaxf(x[i].w, floor_v) + 4.0f) / 4.0f;
          SEGMA_DEV_ASSERT(rank * kRows + row + i < kMels && q < kQuads && q < nvq);
          __stcs(o + (row + i) * kQuads + q, r);
        }
        // the eight lanes of a 128-byte line have their data: drop the line (unless the fp16 tile reads it too);
        // rows are whole lines (nvp % 32 == 0) and belong to one CTA
        if (!out_tm && (q & 7) == 0) {
#pragma unroll
          for (int i = 0; i < kInFlight; ++i) l2_discard_128(src + (row + i) * nvq + q);
        }
      }
    }
    // the constant tail: nothing to read; a thread keeps its columns and walks the rows
    const float4 f4 = make_float4(fill, fill, fill, fill);
    constexpr int kTailIt = (kQuads + kThreadsF - 1) / kThreadsF;
    float4* ocol = o + nq + tid;
#pragma unroll 2
    for (int row = 0; row < kRows; ++row) {
#pragma unroll
      for (int it = 0; it < kTailIt; ++it)
        if (nq + tid + it * kThreadsF < kQuads) __stcs(ocol + it * kThreadsF, f4);
      ocol += kQuads;
    }
  }
  if (out_tm) {
    // fp16 time-major (3002, 80): rows 0 and 3001 are the conv padding; tiles of 32 frames in turn over the cluster
    float (*tile)[kTmTileF + 1] = reinterpret_cast<float (*)[kTmTileF + 1]>(r1);
    __half* o = out_tm + (long long)win * (kFramesOut + 2) * kMels;
    if (rank == 0 && tid < kMels) {
      o[tid] = __float2half(0.f);
      o[(long long)(kFramesOut + 1) * kMels + tid] = __float2half(0.f);
    }
    constexpr int kTiles = (kFramesOut + kTmTileF - 1) / kTmTileF;
    __syncthreads();  // r1 is free (also when no group ran)
#pragma unroll 1
    for (int j = rank; j < kTiles; j += kCl) {
      const int tt = j * kTmTileF;
      const bool live = tt < n_valid;  // uniform per CTA
      if (live) {
        for (int id = tid; id < kMels * kTmTileF; id += kThreadsF) {
          const int m = id / kTmTileF, f = id - m * kTmTileF;
          const int t = tt + f;
          tile[m][f] = (t < n_valid) ? (fmaxf(__ldcg(ls + (long long)m * nvp + t), floor_v) + 4.0f) / 4.0f : fill;
        }
        __syncthreads();
      }
      for (int id = tid; id < kTmTileF * kMels / 2; id += kThreadsF) {
        const int f = id / (kMels / 2), m = (id - f * (kMels / 2)) * 2;
        const int t = tt + f;
        if (t >= kFramesOut) continue;
        float a = fill, c = fill;
        if (live) { a = tile[m][f]; c = tile[m + 1][f]; }
        *reinterpret_cast<uint32_t*>(o + (long long)(t + 1) * kMels + m) = pack_f16x2(a, c);
      }
      if (live) __syncthreads();
    }
  }
}

// Persistent clusters walk the windows.  The finish of window i is deferred until the cluster's transform of window
// i + 1 is done: barrier.cluster.arrive after the transform, barrier.cluster.wait only then, so no CTA ever idles
// waiting for its peers' maxima (the wait was 22 % of all warp time when it followed the arrive directly).
__global__ void __launch_bounds__(kThreadsF, 4) logmel_fused_kernel(const float* __restrict__ pcm, long long pcm_len,
                                                                   int win_len, long long step, int n_windows,
                                                                   int nvp, float* __restrict__ logspec,
                                                                   float* __restrict__ out_f32,
                                                                   __half* __restrict__ out_tm) {
  extern __shared__ __align__(16) unsigned char fused_smem[];
  unsigned char* r1 = fused_smem;                                        // exchange rows / spectrum / transpose tile
  float* stage = reinterpret_cast<float*>(fused_smem + kR1Bytes);
  float2* P2 = reinterpret_cast<float2*>(fused_smem + kR1Bytes + kStageWords * 4);
  float* s_melw = reinterpret_cast<float*>(fused_smem + kR1Bytes + kStageWords * 4 + kPairs * kPP * 8);
  __shared__ float s_red[kThreadsF / 32];
  __shared__ float s_cmax[3];  // maxima of three windows in flight (a peer may be one transform ahead)
  float2* E = reinterpret_cast<float2*>(r1);
  float2* Z = reinterpret_cast<float2*>(r1);
  const int tid = threadIdx.x;
  const uint32_t rank = cluster_ctarank();
  const int n_clusters = gridDim.x / kCl;
  const int p = tid / 20, b = tid - 20 * p;   // pass 1: (pair, b); pass 2: (pair, c = b)

  // the mel taps sit in shared memory (their loads head a dependent chain); the Hann window comes through L1 and the
  // twiddles are computed: the LSU is the busiest pipe of this kernel
  for (int i = tid; i < kMels * kMelTapsSmem; i += kThreadsF) s_melw[i] = g_tab.mel_w2[i / kMelTapsSmem][i % kMelTapsSmem].x;
  // mel filter of this thread in step 4: filter m for the group's pairs 4 h .. 4 h + 3
  const int mel_m = tid >> 1, mel_h = tid & 1;
  const bool wide_bank = __ldg(g_tab.mel_len + mel_m) > kMelTapsSmem;
  const int mel_k0 = __ldg(g_tab.mel_k0 + mel_m), mel_len = __ldg(g_tab.mel_len + mel_m);
  const float2* mel_w = g_tab.mel_w2[mel_m];
  const uint64_t keep_policy = l2_keep_policy();
  const float2 tw_b1 = __ldg(&g_tab.tw_b1[b]), tw_b4 = __ldg(&g_tab.tw_b4[b]);
  __syncthreads();

  int prev_win = -1, slot = 0;
  for (int win = blockIdx.x / kCl; win < n_windows; win += n_clusters) {
    long long avail;
    const int n_valid = window_valid_frames(pcm_len, win, step, win_len, avail);
    const float* w = pcm + (long long)win * step;
    const bool aligned8 = ((reinterpret_cast<uintptr_t>(w) & 7) == 0);
    float* ls = logspec + (long long)win * kMels * nvp;

    // this CTA's share: a contiguous range of frame pairs (a multiple of 4 pairs, so that groups start on 8 frames)
    const int pairs_total = (n_valid + 1) / 2;
    const int per = ((pairs_total + kCl - 1) / kCl + 3) & ~3;
    const int pair_lo = min((int)rank * per, pairs_total), pair_hi = min(pair_lo + per, pairs_total);

    float local_max = -INFINITY;
    if (pair_lo < pair_hi) {
      stage_group(stage, w, (long long)kHop * 2 * pair_lo - 200, avail, aligned8, tid);
      asm volatile("cp.async.wait_all;" ::: "memory");
    }
    __syncthreads();
    // Three barriers per group: after the exchange rows, after the spectrum, after the powers.  Step 4 of a group
    // needs no barrier behind it (it reads only P2 and writes global memory), so warps with short mel filters run
    // ahead into the next group's pass 1 while the warp with the 14-tap filters finishes.
    for (int pair0 = pair_lo; pair0 < pair_hi; pair0 += kPairs) {
      const int np = min(kPairs, pair_hi - pair0);  // live pairs of this group
      const int t0 = 2 * pair0;
      const bool pair_on = p < np;
      // 1. pass 1: window, DFT over a of z[20 a + b], twiddle, store transposed as E[pair][c][b]
      if (pair_on) {
        cpx v[20];
        const float* sa = stage + 2 * kChunkPad * p + b;
#pragma unroll
        for (int a = 0; a < 20; ++a) {
          const int off = 20 * a + 10 * (a >> 3);  // sample 20 a + b of the frame: chunk a / 8 (b < 20)
          const float h = __ldg(g_tab.hann + 20 * a + b);
          v[a] = cpx_make(sa[off] * h, sa[off + kChunkPad] * h);
        }
        dft20(v);
        float2* e = E + p * kEP + b;
        SEGMA_DEV_ASSERT(p < kPairs && b < 20 && (p * kEP + b + 19 * kEL) * 8 < kR1Bytes);
        e[0] = cpx_f2(v[0]);
        // twiddles W400^{b c} = w^c, w = W400^b, from two table values (w, w^4) by at most three further products
        // (c = 4 g + r: w^c = (w^4)^g w^r): the LSU is the busiest pipe of this kernel, the FMA pipes are at 15 %
        const float2 wr[4] = {make_float2(1.f, 0.f), tw_b1, cmul(tw_b1, tw_b1), cmul(cmul(tw_b1, tw_b1), tw_b1)};
        const float2 w8 = cmul(tw_b4, tw_b4);
        const float2 wg[5] = {make_float2(1.f, 0.f), tw_b4, w8, cmul(w8, tw_b4), cmul(w8, w8)};
#pragma unroll
        for (int c = 1; c < 20; ++c) {
          const int g = c >> 2, r = c & 3;
          const float2 tw = g == 0 ? wr[r] : (r == 0 ? wg[g] : cmul(wg[g], wr[r]));
          e[c * kEL] = cmul(cpx_f2(v[c]), tw);
        }
      }
      __syncthreads();
      // every thread has its samples: the next group's may land (in flight until the barrier behind step 3)
      const bool more = pair0 + kPairs < pair_hi;
      const long long s_next = (long long)kHop * (t0 + kGroup) - 200;
      if (more) stage_group(stage, w, s_next, avail, aligned8, tid);
      // 2. pass 2: DFT over b of row c, written back over the same row: Z[pair][c][d] is bin c + 20 d
      if (pair_on) {
        cpx v[20];
        float2* e = E + p * kEP + b * kEL;
#pragma unroll
        for (int i = 0; i < 20; ++i) v[i] = cpx_make(e[i].x, e[i].y);
        dft20(v);
#pragma unroll
        for (int d = 0; d < 20; ++d) e[d] = cpx_f2(v[d]);
      }
      __syncthreads();
      // 3. separate the two frames and take the powers (times 4).  Bin k = c + 20 d sits at row c, column d: sixteen
      //    lanes share d and take c = 0..15 (8-byte bank 5 c + d: conflict-free, also for the mirrored bin); the bins
      //    with c >= 16 and bin 200 make up a second, short round
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int k = half == 0 ? (tid & 15) + 20 * (tid >> 4) : (tid < 40 ? 16 + (tid & 3) + 20 * (tid >> 2) : 200);
        if (half == 0 || tid <= 40) {
          const int k2 = k == 0 ? 0 : kNfft - k;
          const float2* z1p = Z + (k % 20) * kEL + k / 20;
          const float2* z2p = Z + (k2 % 20) * kEL + k2 / 20;
          float2* pw = P2 + k;
          // all loads first, then the stores; pairs beyond np hold stale but finite data
          float2 z1[kPairs], z2[kPairs];
#pragma unroll
          for (int q = 0; q < kPairs; ++q) {
            z1[q] = z1p[q * kEP];
            z2[q] = z2p[q * kEP];
          }
#pragma unroll
          for (int q = 0; q < kPairs; ++q) {
            if (q < np) {
              float sx, sy, dx, dy;
              f2_unpack(f2_add(f2_pack(z1[q].x, z1[q].y), f2_pack(z2[q].x, z2[q].y)), sx, sy);
              f2_unpack(cpx_sub(f2_pack(z1[q].x, z1[q].y), f2_pack(z2[q].x, z2[q].y)), dx, dy);
              pw[q * kPP] = make_float2(fmaf(sx, sx, dy * dy), fmaf(dx, dx, sy * sy));  // |Z1 + conj Z2|^2, |Z1 - conj Z2|^2
            }
          }
        }
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncthreads();
      // 4. sparse mel filters on (P_A, P_B) pairs + log10: each tap is loaded once for four frame pairs; warps hold
      //    filters of similar length (the slaney bank widens with m)
      if (4 * mel_h < np) {
        uint64_t acc[4] = {0ull, 0ull, 0ull, 0ull};  // (0.f, 0.f)
        const float2* pq = P2 + 4 * mel_h * kPP + mel_k0;
#pragma unroll 2
        for (int i = 0; i < mel_len; ++i) {
          const float wsc = wide_bank ? __ldg(mel_w + i).x : s_melw[mel_m * kMelTapsSmem + i];
          const float2 wv = make_float2(wsc, wsc);
          const uint64_t w2 = f2_pack(wv.x, wv.y);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 pv = pq[j * kPP + i];
            acc[j] = f2_fma(w2, f2_pack(pv.x, pv.y), acc[j]);
          }
        }
        float* dst = ls + (long long)mel_m * nvp + t0 + 8 * mel_h;
        float lv[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float ma, mb;
          f2_unpack(acc[j], ma, mb);
          lv[2 * j] = 0.30102999566398120f * __log2f(fmaxf(ma, 1e-10f));  // log10
          lv[2 * j + 1] = 0.30102999566398120f * __log2f(fmaxf(mb, 1e-10f));
        }
        SEGMA_DEV_ASSERT(win < n_windows && t0 + 8 * mel_h + 7 < nvp);
        // a window's last pair may end one frame past n_valid: that frame is all zeros (-10, which the maximum
        // contains anyway whenever n_valid < 3000) and the finish never reads it
        if (4 * mel_h + 3 < np) {
          // the thread's eight frames are one 32-byte sector of the row: one 256-bit store
          st_f32x8_keep(dst, lv, keep_policy);
#pragma unroll
          for (int j = 0; j < 8; ++j) local_max = fmaxf(local_max, lv[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (4 * mel_h + j < np) {
              st_f32x2_keep(dst + 2 * j, lv[2 * j], lv[2 * j + 1], keep_policy);
              local_max = fmaxf(local_max, fmaxf(lv[2 * j], lv[2 * j + 1]));
            }
          }
        }
      }
    }

    // ---- this CTA's maximum; the finish of the previous window; hand this window to the cluster --------------
    local_max = warp_max(local_max);
    if (lane_id() == 0) s_red[tid >> 5] = local_max;
    __syncthreads();
    if (tid == 0) {
      float m = s_red[0];
      for (int i = 1; i < kThreadsF / 32; ++i) m = fmaxf(m, s_red[i]);
      s_cmax[slot] = m;
    }
    __syncthreads();  // s_cmax[slot] is written; the exchange rows are free for the transpose tile
    float gmax_prev = -INFINITY;
    if (prev_win >= 0) {
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");  // peers finished the previous transform
      const int pslot = slot == 0 ? 2 : slot - 1;
#pragma unroll
      for (uint32_t r = 0; r < kCl; ++r) gmax_prev = fmaxf(gmax_prev, ld_cluster_f32(&s_cmax[pslot], r));
    }
    // hand this window to the cluster before the stores of the previous one's finish are issued (a release behind
    // them would wait for them to drain); the maxima read above are in registers
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    if (prev_win >= 0) {
      long long pa;
      const int pv = window_valid_frames(pcm_len, prev_win, step, win_len, pa);
      finish_window(prev_win, rank, tid, pv, gmax_prev, nvp, logspec + (long long)prev_win * kMels * nvp, out_f32, out_tm, r1);
    }
    prev_win = win;
    slot = slot == 2 ? 0 : slot + 1;
  }
  if (prev_win >= 0) {
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    const int pslot = slot == 0 ? 2 : slot - 1;
    float gmax = -INFINITY;
#pragma unroll
    for (uint32_t r = 0; r < kCl; ++r) gmax = fmaxf(gmax, ld_cluster_f32(&s_cmax[pslot], r));
    long long pa;
    const int pv = window_valid_frames(pcm_len, prev_win, step, win_len, pa);
    finish_window(prev_win, rank, tid, pv, gmax, nvp, logspec + (long long)prev_win * kMels * nvp, out_f32, out_tm, r1);
  }
  cluster_sync_all();  // no CTA leaves while a peer may still read its maximum
}

static int max_valid_frames(int win_len) {
  long long nv = ((long long)win_len + 200 + kHop - 1) / kHop;
  if (nv > kFramesOut) nv = kFramesOut;
  return (int)nv;
}
// scratch row length: a multiple of 32 floats, so that every row starts on a 128-byte line (the finish drops whole
// lines from L2 after reading them: a line shared by two rows would lose the other row's values)
static int padded_valid(int win_len) { return ceil_div(max_valid_frames(win_len), 32) * 32; }

}  // namespace segma

using namespace segma;

extern "C" {

size_t segma_logmel_scratch_bytes(int n_windows, int win_len) {
  if (n_windows <= 0 || win_len <= 0) return 0;
  const size_t head = ((size_t)n_windows * sizeof(uint32_t) + 255) / 256 * 256;
  return head + (size_t)n_windows * kMels * padded_valid(win_len) * sizeof(float);
}

int segma_logmel_set_filters(const float* mel_201x80) {
  SEGMA_REQUIRE(mel_201x80 != nullptr, "segma_logmel_set_filters: NULL matrix");
  std::lock_guard<std::mutex> lock(g_tab_mutex);
  std::vector<float> previous = g_mel_dense;
  g_mel_dense.assign(mel_201x80, mel_201x80 + (size_t)kBins * kMels);
  for (bool& ready : g_tab_ready.done) ready = false;  // other devices pick the new filterbank up on their next call
  const int rc = upload_tables_locked();
  if (rc != SEGMA_OK) g_mel_dense = previous;  // a rejected bank leaves the active one in place (re-uploaded on the next call)
  return rc;
}

int segma_logmel_get_filters(float* mel_201x80) {
  SEGMA_REQUIRE(mel_201x80 != nullptr, "segma_logmel_get_filters: NULL matrix");
  std::lock_guard<std::mutex> lock(g_tab_mutex);
  if (g_mel_dense.empty()) default_mel(g_mel_dense);
  std::memcpy(mel_201x80, g_mel_dense.data(), sizeof(float) * kBins * kMels);
  return SEGMA_OK;
}

int segma_logmel(const float* pcm, int64_t pcm_len, int n_windows, int win_len, int64_t step, float* out_f32,
                 void* out_tm, void* scratch, void* stream) {
  SEGMA_REQUIRE(n_windows >= 0, "segma_logmel: negative n_windows");
  if (n_windows == 0) return SEGMA_OK;
  SEGMA_REQUIRE(pcm && scratch, "segma_logmel: NULL pcm/scratch");
  SEGMA_REQUIRE(win_len > 0 && win_len <= kPadSamples - kNfft, "segma_logmel: win_len %d outside (0, %d]", win_len,
                kPadSamples - kNfft);
  SEGMA_REQUIRE(step >= 0 && pcm_len >= 0, "segma_logmel: negative step/pcm_len");
  SEGMA_REQUIRE(out_f32 || out_tm, "segma_logmel: no output requested");
  SEGMA_REQUIRE(n_windows <= 65535, "segma_logmel: at most 65535 windows per call");
  SEGMA_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 127) == 0, "segma_logmel: scratch must be 128-byte aligned");
  int rc = ensure_tables();
  if (rc != SEGMA_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int nvp = padded_valid(win_len);
  const size_t head = ((size_t)n_windows * sizeof(uint32_t) + 255) / 256 * 256;
  float* logspec = reinterpret_cast<float*>(static_cast<char*>(scratch) + head);
  {
    static PerDeviceFlag fused_attr;
    static int max_clusters[kMaxDevices];
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(kThreadsF);
    cfg.dynamicSmemBytes = kFusedSmem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int dev = current_device();
    if (!fused_attr.here()) {
      SEGMA_CUDA_OK(cudaFuncSetAttribute(logmel_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmem));
      cfg.gridDim = dim3(kCl * 4 * device_sm_count());
      int n = 0;
      SEGMA_CUDA_OK(cudaOccupancyMaxActiveClusters(&n, logmel_fused_kernel, &cfg));
      SEGMA_REQUIRE(n > 0, "segma_logmel: no cluster of %d CTAs fits on this device", kCl);
      max_clusters[dev] = n;
      fused_attr.here() = true;
    }
    cfg.gridDim = dim3((unsigned)std::min(n_windows, max_clusters[dev]) * kCl);  // persistent: every cluster is resident
    SEGMA_CUDA_OK(cudaLaunchKernelEx(&cfg, logmel_fused_kernel, pcm, (long long)pcm_len, win_len, (long long)step,
                                     n_windows, nvp, logspec, out_f32, static_cast<__half*>(out_tm)));
    return launch_status("logmel_fused_kernel");
  }
  return SEGMA_OK;
}

}  // extern "C"
