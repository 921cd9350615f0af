// LSTM recurrence over the window axis and the per-label linear heads.
//
// The reference builds nn.LSTM without batch_first (src/segma/models/whisper/hydra.py:48-51,
// surgical_hydra.py:57-60), so the (B, T, d) encoder output is consumed as (seq = B windows, batch = T
// frames): the recurrence runs across the <= batch_size windows of one forward call and the kept frames
// are independent rows (SURVEY.md finding 6).  The input projection x W_ih^T + b is a tcgen05 GEMM
// (gemm_tc5.cu) on a split-precision operand pair (segma_cast_f16_split); this kernel runs the sequential part:
// each CTA owns R frames of one direction, keeps h in shared memory and c in registers, and walks the n_steps
// windows.  fp32 throughout (W_hh at 22+ bits, see lstm_layer_smem_kernel).
#include <type_traits>

#include "common.cuh"

namespace segma {

constexpr int kLstmRows = 8;  // frames per CTA

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int H>
__global__ void __launch_bounds__(4 * H) lstm_layer_kernel(const float* __restrict__ pre,
                                                            const float* __restrict__ w_hh_t, int n_steps,
                                                            int n_rows, int n_dirs, float* __restrict__ out,
                                                            __half* __restrict__ out_f16) {
  constexpr int R = kLstmRows;
  constexpr int G = 4 * H;
  constexpr int kItems = (R * H) / G;  // pointwise items per thread
  __shared__ float s_h[R][H];
  __shared__ float s_g[R][G];
  const int dir = blockIdx.y;
  const int r0 = blockIdx.x * R;
  const int j = threadIdx.x;
  const float* w = w_hh_t + (long long)dir * H * G + j;
  float c_state[kItems];
#pragma unroll
  for (int i = 0; i < kItems; ++i) c_state[i] = 0.f;
  for (int i = j; i < R * H; i += G) (&s_h[0][0])[i] = 0.f;
  __syncthreads();

  for (int step = 0; step < n_steps; ++step) {
    const int s = dir == 0 ? step : n_steps - 1 - step;
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = r0 + r;
      acc[r] = row < n_rows ? __ldg(pre + ((long long)s * n_rows + row) * (n_dirs * G) + dir * G + j) : 0.f;
    }
#pragma unroll 8
    for (int k = 0; k < H; ++k) {
      const float wk = __ldg(w + (long long)k * G);
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = fmaf(wk, s_h[r][k], acc[r]);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) s_g[r][j] = acc[r];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
      const int idx = j + i * G;
      const int r = idx / H, u = idx - r * H;
      const float ig = sigmoid_acc(s_g[r][u]);
      const float fg = sigmoid_acc(s_g[r][H + u]);
      const float gg = tanhf(s_g[r][2 * H + u]);
      const float og = sigmoid_acc(s_g[r][3 * H + u]);
      const float c = fmaf(fg, c_state[i], ig * gg);
      c_state[i] = c;
      const float h = og * tanhf(c);
      s_h[r][u] = h;
      const int row = r0 + r;
      if (row < n_rows) {
        const long long o = ((long long)s * n_rows + row) * (n_dirs * H) + dir * H + u;
        out[o] = h;
        if (out_f16) out_f16[o] = __float2half(h);
      }
    }
    __syncthreads();
  }
}

// Same recurrence with W_hh resident on the SM, R = 3 frames per CTA so that 2 x ceil(199 / 3) = 134 CTAs cover the
// 148 SMs in one wave.  Per step a thread owns one gate column: H x (weight fetch + R FMAs), then the pointwise cell
// update; the fetch of W_hh no longer depends on L2 latency, which bounded the kernel above at ~11 us per step.
// The weights keep fp32 precision -- the recurrence runs over up to batch_size = 128 windows, and a rounded W_hh is a
// systematic perturbation that adds up along it:
//   H = 64  : fp32 in shared memory (64 KB)
//   H = 128 : fp32 does not fit (256 KB), so w = hi + lo * 2^-11 with hi = fp16(w) in shared memory (128 KB) and
//             lo = fp16((w - hi) * 2^11) in 64 registers of the thread that owns the column (22 significant bits)
constexpr int kLstmRowsSmem = 3;
constexpr float kLoScale = 2048.0f;

template <int H>
__global__ void __launch_bounds__(4 * H) lstm_layer_smem_kernel(const float* __restrict__ pre,
                                                                 const float* __restrict__ w_hh_t, int n_steps,
                                                                 int n_rows, int n_dirs, float* __restrict__ out,
                                                                 __half* __restrict__ out_f16) {
  constexpr int R = kLstmRowsSmem;
  constexpr int G = 4 * H;
  constexpr bool kSplit = H > 64;
  using WT = typename std::conditional<kSplit, __half, float>::type;
  extern __shared__ __align__(16) unsigned char lstm_smem[];
  WT* s_w = reinterpret_cast<WT*>(lstm_smem);                            // [H][G]
  float* s_h = reinterpret_cast<float*>(lstm_smem + sizeof(WT) * H * G);  // [H][4] (R padded to 4)
  float* s_g = s_h + H * 4;                                               // [R][G]
  const int dir = blockIdx.y;
  const int r0 = blockIdx.x * R;
  const int j = threadIdx.x;
  __half2 w_lo[kSplit ? H / 2 : 1];
  if (kSplit) {
#pragma unroll
    for (int k = 0; k < H; k += 2) {
      const float w0 = __ldg(w_hh_t + ((long long)dir * H + k) * G + j);
      const float w1 = __ldg(w_hh_t + ((long long)dir * H + k + 1) * G + j);
      const __half h0 = __float2half(w0), h1 = __float2half(w1);
      s_w[k * G + j] = h0;
      s_w[(k + 1) * G + j] = h1;
      w_lo[k / 2] = __floats2half2_rn((w0 - __half2float(h0)) * kLoScale, (w1 - __half2float(h1)) * kLoScale);
    }
  } else {
    for (int k = 0; k < H; ++k) s_w[k * G + j] = __ldg(w_hh_t + ((long long)dir * H + k) * G + j);
  }
  for (int i = j; i < H * 4; i += G) s_h[i] = 0.f;
  float c_state = 0.f;
  __syncthreads();
  const int pr = j / H, pu = j - pr * H;  // pointwise item of this thread (valid if pr < R)
  // the input pre-activations of step t + 1 are fetched while step t runs: their HBM / L2 latency (about a
  // microsecond, 128 times per launch) would otherwise sit at the head of every step's dependency chain
  float nxt[R];
  {
    const int s0 = dir == 0 ? 0 : n_steps - 1;
#pragma unroll
    for (int r = 0; r < R; ++r)
      nxt[r] = r0 + r < n_rows ? __ldg(pre + ((long long)s0 * n_rows + r0 + r) * (n_dirs * G) + dir * G + j) : 0.f;
  }
  for (int step = 0; step < n_steps; ++step) {
    const int s = dir == 0 ? step : n_steps - 1 - step;
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = nxt[r];
    if (step + 1 < n_steps) {
      const int s1 = dir == 0 ? step + 1 : n_steps - 2 - step;
#pragma unroll
      for (int r = 0; r < R; ++r)
        nxt[r] = r0 + r < n_rows ? __ldg(pre + ((long long)s1 * n_rows + r0 + r) * (n_dirs * G) + dir * G + j) : 0.f;
    }
    if (kSplit) {
#pragma unroll
      for (int k = 0; k < H; k += 2) {
        const float2 lo = __half22float2(w_lo[k / 2]);
        const float wa = fmaf(lo.x, 1.0f / kLoScale, __half2float(s_w[k * G + j]));
        const float wb = fmaf(lo.y, 1.0f / kLoScale, __half2float(s_w[(k + 1) * G + j]));
        const float4 ha = *reinterpret_cast<const float4*>(s_h + k * 4);
        const float4 hb = *reinterpret_cast<const float4*>(s_h + k * 4 + 4);
        acc[0] = fmaf(wa, ha.x, acc[0]);
        acc[1] = fmaf(wa, ha.y, acc[1]);
        acc[2] = fmaf(wa, ha.z, acc[2]);
        acc[0] = fmaf(wb, hb.x, acc[0]);
        acc[1] = fmaf(wb, hb.y, acc[1]);
        acc[2] = fmaf(wb, hb.z, acc[2]);
      }
    } else {
#pragma unroll 16
      for (int k = 0; k < H; ++k) {
        const float wk = s_w[k * G + j];
        const float4 hv = *reinterpret_cast<const float4*>(s_h + k * 4);
        acc[0] = fmaf(wk, hv.x, acc[0]);
        acc[1] = fmaf(wk, hv.y, acc[1]);
        acc[2] = fmaf(wk, hv.z, acc[2]);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) s_g[r * G + j] = acc[r];
    __syncthreads();
    if (pr < R) {
      const float ig = sigmoid_acc(s_g[pr * G + pu]);
      const float fg = sigmoid_acc(s_g[pr * G + H + pu]);
      const float gg = tanhf(s_g[pr * G + 2 * H + pu]);
      const float og = sigmoid_acc(s_g[pr * G + 3 * H + pu]);
      c_state = fmaf(fg, c_state, ig * gg);
      const float h = og * tanhf(c_state);
      s_h[pu * 4 + pr] = h;
      const int row = r0 + pr;
      if (row < n_rows) {
        SEGMA_DEV_ASSERT(s >= 0 && s < n_steps && pu < H && isfinite(h));
        const long long o = ((long long)s * n_rows + row) * (n_dirs * H) + dir * H + pu;
        out[o] = h;
        if (out_f16) out_f16[o] = __float2half(h);
      }
    }
    __syncthreads();
  }
}

// one warp per (step, frame) row: C dot products of length n_feat
__global__ void __launch_bounds__(256) heads_kernel(const float* __restrict__ feat, int n_steps, int n_rows,
                                                     int n_feat, int n_keep, const float* __restrict__ w,
                                                     const float* __restrict__ b, int C, float* __restrict__ logits,
                                                     long long frame_offset, int step_frames,
                                                     const long long* __restrict__ frame_offsets) {
  const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= (long long)n_steps * n_keep) return;
  const int s = (int)(wid / n_keep), r = (int)(wid - (long long)s * n_keep);
  const int lane = lane_id();
  const float* f = feat + ((long long)s * n_rows + r) * n_feat;
  // window s writes frames frame_offset + s * step_frames + r, or frame_offsets[s] + r for windows packed from several files
  float* dst = logits + ((frame_offsets ? frame_offsets[s] : frame_offset + (long long)s * step_frames) + r) * C;
  for (int c = 0; c < C; ++c) {
    float acc = 0.f;
    for (int k = lane; k < n_feat; k += 32) acc = fmaf(__ldg(f + k), __ldg(w + (long long)c * n_feat + k), acc);
    acc = warp_sum(acc);
    if (lane == 0) dst[c] = acc + __ldg(b + c);
  }
}

}  // namespace segma

using namespace segma;

extern "C" {

int segma_lstm_layer(const float* pre, const float* w_hh_t, int n_steps, int n_rows, int hidden, int n_dirs,
                     float* out, void* out_f16, void* stream) {
  SEGMA_REQUIRE(n_steps >= 0 && n_rows >= 0 && (n_dirs == 1 || n_dirs == 2), "segma_lstm_layer: bad shape");
  if (n_steps == 0 || n_rows == 0) return SEGMA_OK;
  SEGMA_REQUIRE(pre && w_hh_t && out, "segma_lstm_layer: NULL buffer");
  dim3 grid(ceil_div(n_rows, kLstmRows), n_dirs);
  cudaStream_t st = (cudaStream_t)stream;
  __half* ob = static_cast<__half*>(out_f16);
  dim3 grid_s(ceil_div(n_rows, kLstmRowsSmem), n_dirs);
  switch (hidden) {
    case 64: {
      constexpr int kSmem = 64 * 256 * 4 + 64 * 4 * 4 + kLstmRowsSmem * 256 * 4;  // fp32 weights
      static PerDeviceFlag attr_set;
      if (!attr_set.here()) {
        SEGMA_CUDA_OK(cudaFuncSetAttribute(lstm_layer_smem_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        attr_set.here() = true;
      }
      lstm_layer_smem_kernel<64><<<grid_s, 256, kSmem, st>>>(pre, w_hh_t, n_steps, n_rows, n_dirs, out, ob);
      break;
    }
    case 128: {
      constexpr int kSmem = 128 * 512 * 2 + 128 * 4 * 4 + kLstmRowsSmem * 512 * 4;
      static PerDeviceFlag attr_set;
      if (!attr_set.here()) {
        SEGMA_CUDA_OK(cudaFuncSetAttribute(lstm_layer_smem_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        attr_set.here() = true;
      }
      lstm_layer_smem_kernel<128><<<grid_s, 512, kSmem, st>>>(pre, w_hh_t, n_steps, n_rows, n_dirs, out, ob);
      break;
    }
    // H = 256: W_hh (1 MB per direction in fp32) stays in L2, read with __ldg every step
    case 256: lstm_layer_kernel<256><<<grid, 1024, 0, st>>>(pre, w_hh_t, n_steps, n_rows, n_dirs, out, ob); break;
    default:
      set_last_error("segma_lstm_layer: hidden size %d not supported (64, 128, 256)", hidden);
      return SEGMA_ERR_UNSUPPORTED;
  }
  return launch_status("lstm_layer_kernel");
}

int segma_heads(const float* feat, int n_steps, int n_rows, int n_feat, int n_keep, const float* w, const float* b,
                int n_labels, float* logits, int64_t frame_offset, int step_frames, void* stream) {
  SEGMA_REQUIRE(n_steps >= 0 && n_rows > 0 && n_feat > 0 && n_keep >= 0 && n_keep <= n_rows && n_labels > 0,
                "segma_heads: bad shape");
  if (n_steps == 0 || n_keep == 0) return SEGMA_OK;
  SEGMA_REQUIRE(feat && w && b && logits, "segma_heads: NULL buffer");
  const long long warps = (long long)n_steps * n_keep;
  heads_kernel<<<(unsigned)ceil_div_ll(warps, 8), 256, 0, (cudaStream_t)stream>>>(
      feat, n_steps, n_rows, n_feat, n_keep, w, b, n_labels, logits, frame_offset, step_frames, nullptr);
  return launch_status("heads_kernel");
}

int segma_heads_at(const float* feat, int n_steps, int n_rows, int n_feat, int n_keep, const float* w, const float* b,
                   int n_labels, float* logits, const int64_t* frame_offsets, void* stream) {
  SEGMA_REQUIRE(n_steps >= 0 && n_rows > 0 && n_feat > 0 && n_keep >= 0 && n_keep <= n_rows && n_labels > 0,
                "segma_heads_at: bad shape");
  if (n_steps == 0 || n_keep == 0) return SEGMA_OK;
  SEGMA_REQUIRE(feat && w && b && logits && frame_offsets, "segma_heads_at: NULL buffer");
  const long long warps = (long long)n_steps * n_keep;
  heads_kernel<<<(unsigned)ceil_div_ll(warps, 8), 256, 0, (cudaStream_t)stream>>>(
      feat, n_steps, n_rows, n_feat, n_keep, w, b, n_labels, logits, 0, 0,
      reinterpret_cast<const long long*>(frame_offsets));
  return launch_status("heads_kernel");
}

}  // extern "C"
