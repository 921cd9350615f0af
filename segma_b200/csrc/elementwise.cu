// Row-wise LayerNorm (fp32 statistics) with the layer-weighted hidden-state mix folded in, and a strided
// fp32 -> fp16 cast.  Replaces ATen LayerNorm inside WhisperEncoderLayer / torchaudio EncoderLayer
// (site-packages/transformers/models/whisper/modeling_whisper.py:380-412, 643;
// site-packages/torchaudio/models/wav2vec2/components.py:363-401) and torch.stack + einsum of
// src/segma/models/whisper/surgical_hydra.py:82-98.  HBM-bound: one read of the fp32 residual row,
// one fp16 (and/or fp32) write, 128-bit accesses, one warp per row.
#include <algorithm>

#include "common.cuh"

namespace segma {

constexpr int kLnWarps = 8;

template <int VPL>  // float4 chunks per lane: d = VPL * 128
__global__ void __launch_bounds__(kLnWarps * 32) layernorm_kernel(
    const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, long long rows,
    __half* __restrict__ out_f16, float* __restrict__ out_f32, float* __restrict__ mix, int period,
    int n_keep, float w_in, float w_out, int mix_init, int only_kept) {
  constexpr int d = VPL * 128;
  long long row = (long long)blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  if (only_kept) {  // enumerate kept rows only: row = window * period + frame, frame < n_keep
    if (row >= (rows / period) * n_keep) return;
    row = (row / n_keep) * period + row % n_keep;
  } else if (row >= rows) {
    return;
  }
  const int lane = lane_id();
  const float4* xr = reinterpret_cast<const float4*>(x + row * d);
  float4 v[VPL];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    v[i] = xr[lane + 32 * i];
    sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(sum) * (1.0f / d);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
    sq += (a * a + b * b) + (c * c + e * e);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(sq) * (1.0f / d) + 1e-5f);
  const bool do_mix = mix != nullptr && (row % period) < n_keep;
  float4* mrow = nullptr;
  if (do_mix) mrow = reinterpret_cast<float4*>(mix + ((row / period) * n_keep + (row % period)) * d);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c4 = lane + 32 * i;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
    float4 y;
    y.x = (v[i].x - mean) * rstd * g.x + b.x;
    y.y = (v[i].y - mean) * rstd * g.y + b.y;
    y.z = (v[i].z - mean) * rstd * g.z + b.z;
    y.w = (v[i].w - mean) * rstd * g.w + b.w;
    if (out_f16) {
      uint2 p;
      p.x = pack_f16x2(y.x, y.y);
      p.y = pack_f16x2(y.z, y.w);
      reinterpret_cast<uint2*>(out_f16 + row * d)[c4] = p;
    }
    if (out_f32) reinterpret_cast<float4*>(out_f32 + row * d)[c4] = y;
    if (do_mix) {
      float4 m = mix_init ? make_float4(0.f, 0.f, 0.f, 0.f) : mrow[c4];
      m.x += w_in * v[i].x + w_out * y.x;
      m.y += w_in * v[i].y + w_out * y.y;
      m.z += w_in * v[i].z + w_out * y.z;
      m.w += w_in * v[i].w + w_out * y.w;
      mrow[c4] = m;
    }
  }
}

__global__ void __launch_bounds__(256) cast_f16_kernel(const float* __restrict__ src, long long lds,
                                                         __half* __restrict__ dst, long long ldd,
                                                         long long rows, int cols) {
  const int quads = cols / 4;
  const long long total = rows * quads;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / quads;
    const int q = (int)(i - r * quads);
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + r * lds) + q);
    uint2 p;
    p.x = pack_f16x2(v.x, v.y);
    p.y = pack_f16x2(v.z, v.w);
    reinterpret_cast<uint2*>(dst + r * ldd)[q] = p;
  }
}

// fp32 -> [hi | lo | hi] fp16 with hi = fp16(x), lo = fp16(x - hi): the A operand of a split-precision GEMM whose
// weight is laid out [W_hi | W_hi | W_lo], so that x W^T = hi W_hi + lo W_hi + hi W_lo carries ~22 bits of both
// operands through the fp16 tensor cores (the dropped lo W_lo term is 2^-22 relative).
__global__ void __launch_bounds__(256) cast_f16_split_kernel(const float* __restrict__ src, long long lds,
                                                               __half* __restrict__ dst, long long ldd,
                                                               long long rows, int cols) {
  const int quads = cols / 4;
  const long long total = rows * quads;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / quads;
    const int q = (int)(i - r * quads);
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + r * lds) + q);
    const __half hx = __float2half(v.x), hy = __float2half(v.y), hz = __float2half(v.z), hw = __float2half(v.w);
    uint2 hi, lo;
    hi.x = pack_f16x2(__half2float(hx), __half2float(hy));
    hi.y = pack_f16x2(__half2float(hz), __half2float(hw));
    lo.x = pack_f16x2(v.x - __half2float(hx), v.y - __half2float(hy));
    lo.y = pack_f16x2(v.z - __half2float(hz), v.w - __half2float(hw));
    __half* row = dst + r * ldd;
    reinterpret_cast<uint2*>(row)[q] = hi;
    reinterpret_cast<uint2*>(row + cols)[q] = lo;
    reinterpret_cast<uint2*>(row + 2 * cols)[q] = hi;
  }
}

// integer PCM -> float32 in [-1, 1) exactly as the host decode does (x / 2^15, x / 2^31): the file's samples
// cross PCIe in their native width and are widened on the device
template <typename T>
__global__ void __launch_bounds__(256) pcm_to_f32_kernel(const T* __restrict__ src, long long n, float scale,
                                                          float* __restrict__ dst) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = static_cast<float>(src[i]) * scale;
}

template <int VPL>
static int launch_ln(const float* x, const float* gamma, const float* beta, long long rows, void* out_f16,
                     float* out_f32, float* mix, int period, int n_keep, float w_in, float w_out, int mix_init,
                     int only_kept, cudaStream_t st) {
  const long long blocks = ceil_div_ll(only_kept ? (rows / period) * n_keep : rows, kLnWarps);
  layernorm_kernel<VPL><<<(unsigned)blocks, kLnWarps * 32, 0, st>>>(
      x, gamma, beta, rows, static_cast<__half*>(out_f16), out_f32, mix, period, n_keep, w_in, w_out,
      mix_init, only_kept);
  return launch_status("layernorm_kernel");
}

}  // namespace segma

using namespace segma;

extern "C" {

int segma_layernorm(const float* x, const float* gamma, const float* beta, int64_t rows, int d, void* out_f16,
                    float* out_f32, float* mix, int period, int n_keep, float w_in, float w_out, int mix_init,
                    int only_kept, void* stream) {
  SEGMA_REQUIRE(rows >= 0 && d > 0, "segma_layernorm: bad shape");
  if (rows == 0) return SEGMA_OK;
  SEGMA_REQUIRE(x && gamma && beta, "segma_layernorm: NULL input");
  SEGMA_REQUIRE(d % 128 == 0 && d <= 2048, "segma_layernorm: d=%d must be a multiple of 128 and <= 2048", d);
  SEGMA_REQUIRE((mix == nullptr && !only_kept) || (period > 0 && n_keep > 0 && n_keep <= period && rows % period == 0),
                "segma_layernorm: bad period / n_keep geometry");
  SEGMA_REQUIRE(rows < (1ll << 31) * kLnWarps, "segma_layernorm: too many rows");
  cudaStream_t st = (cudaStream_t)stream;
  if (period <= 0) period = 1;
#define SEGMA_LN_CASE(V)                                                                                       \
  case V:                                                                                                      \
    return launch_ln<V>(x, gamma, beta, rows, out_f16, out_f32, mix, period, n_keep, w_in, w_out, mix_init, only_kept, st);
  switch (d / 128) {
    SEGMA_LN_CASE(1) SEGMA_LN_CASE(2) SEGMA_LN_CASE(3) SEGMA_LN_CASE(4) SEGMA_LN_CASE(5) SEGMA_LN_CASE(6)
    SEGMA_LN_CASE(7) SEGMA_LN_CASE(8) SEGMA_LN_CASE(9) SEGMA_LN_CASE(10) SEGMA_LN_CASE(11) SEGMA_LN_CASE(12)
    SEGMA_LN_CASE(13) SEGMA_LN_CASE(14) SEGMA_LN_CASE(15) SEGMA_LN_CASE(16)
  }
#undef SEGMA_LN_CASE
  set_last_error("segma_layernorm: unsupported d=%d", d);
  return SEGMA_ERR_UNSUPPORTED;
}

int segma_pcm_to_f32(const void* src, int format, int64_t n, float* dst, void* stream) {
  SEGMA_REQUIRE(n >= 0, "segma_pcm_to_f32: negative length");
  if (n == 0) return SEGMA_OK;
  SEGMA_REQUIRE(src && dst, "segma_pcm_to_f32: NULL buffer");
  const int grid = (int)std::min<long long>(ceil_div_ll(n, 256), (long long)device_sm_count() * 16);
  cudaStream_t st = (cudaStream_t)stream;
  switch (format) {
    case SEGMA_PCM_S16:
      pcm_to_f32_kernel<short><<<grid, 256, 0, st>>>(static_cast<const short*>(src), n, 1.0f / 32768.0f, dst);
      break;
    case SEGMA_PCM_S32:
      pcm_to_f32_kernel<int><<<grid, 256, 0, st>>>(static_cast<const int*>(src), n, 1.0f / 2147483648.0f, dst);
      break;
    case SEGMA_PCM_F32:
      return check_cuda(cudaMemcpyAsync(dst, src, sizeof(float) * (size_t)n, cudaMemcpyDeviceToDevice, st), "pcm copy");
    default:
      set_last_error("segma_pcm_to_f32: unknown sample format %d", format);
      return SEGMA_ERR_INVALID_ARGUMENT;
  }
  return launch_status("pcm_to_f32_kernel");
}

int segma_cast_f16(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int cols, void* stream) {
  SEGMA_REQUIRE(rows >= 0 && cols > 0, "segma_cast_f16: bad shape");
  if (rows == 0) return SEGMA_OK;
  SEGMA_REQUIRE(src && dst, "segma_cast_f16: NULL buffer");
  SEGMA_REQUIRE(cols % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0, "segma_cast_f16: cols and strides must be multiples of 4");
  const long long total = rows * (cols / 4);
  const int grid = (int)std::min<long long>(ceil_div_ll(total, 256), (long long)device_sm_count() * 16);
  cast_f16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, lds, static_cast<__half*>(dst), ldd, rows, cols);
  return launch_status("cast_f16_kernel");
}

int segma_cast_f16_split(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int cols, void* stream) {
  SEGMA_REQUIRE(rows >= 0 && cols > 0, "segma_cast_f16_split: bad shape");
  if (rows == 0) return SEGMA_OK;
  SEGMA_REQUIRE(src && dst, "segma_cast_f16_split: NULL buffer");
  SEGMA_REQUIRE(cols % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0 && ldd >= 3 * (int64_t)cols,
                "segma_cast_f16_split: cols and strides must be multiples of 4, ldd >= 3 * cols");
  const long long total = rows * (cols / 4);
  const int grid = (int)std::min<long long>(ceil_div_ll(total, 256), (long long)device_sm_count() * 16);
  cast_f16_split_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, lds, static_cast<__half*>(dst), ldd, rows, cols);
  return launch_status("cast_f16_split_kernel");
}

}  // extern "C"
