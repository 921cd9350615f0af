// Fused multi-head self-attention softmax(Q K^T + bias) V, head_dim 64, fp16 in / fp32 softmax / fp16 out.
//
// Replaces torch SDPA as dispatched by WhisperAttention (site-packages/transformers/models/whisper/
// modeling_whisper.py:338-352; T = 1500, no mask) and torchaudio SelfAttention / WavLMSelfAttention
// (site-packages/torchaudio/models/wav2vec2/components.py:305-307, wavlm_attention.py:166-211; T = 199,
// optional gated relative-position bias passed to SDPA as attn_mask).
//
// Flash-style single pass: a CTA owns 64 query rows of one (window, head); K/V tiles of 64 keys stream
// through a double-buffered cp.async ring in XOR-swizzled shared memory; scores never leave registers.
// This version uses the warp-level mma.sync tensor path (m16n8k16 fp16); the tcgen05/TMEM pipeline is the
// GEMM kernel's and is the planned upgrade for this kernel.
#include <cstdlib>

#include "common.cuh"

namespace segma {

constexpr int kHd = 64;        // head dim
constexpr int kQTile = 64;     // queries per CTA (16 per warp)
constexpr int kKTile = 64;     // keys per pipeline stage
constexpr int kAttnThreads = 128;

// 64 x 64 fp16 tile, rows of 128 B split in 8 chunks of 16 B; chunk index XOR (row & 7)
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// load a (rows x 64) fp16 tile from global rows [row0, row0+64) of a matrix with leading dimension ld
__device__ __forceinline__ void load_tile_async(unsigned char* smem_tile, const __half* __restrict__ g,
                                                long long ld, int row0, int row_limit, int tid) {
#pragma unroll
  for (int i = 0; i < (64 * 8) / kAttnThreads; ++i) {
    const int idx = tid + i * kAttnThreads;
    const int r = idx >> 3, c = idx & 7;
    const bool ok = (row0 + r) < row_limit;
    const __half* src = g + (long long)(ok ? row0 + r : 0) * ld + c * 8;
    cp_async_16(smem_tile + tile_off(r, c), src, ok);
  }
}

__global__ void __launch_bounds__(kAttnThreads) attention_kernel(const __half* __restrict__ qkv, int T,
                                                                 int n_heads, int n_query,
                                                                 const float* __restrict__ gate,
                                                                 const float* __restrict__ pos_bias, int pb_ld,
                                                                 __half* __restrict__ out) {
  __shared__ __align__(128) unsigned char s_q[kQTile * 128];
  __shared__ __align__(128) unsigned char s_k[2][kKTile * 128];
  __shared__ __align__(128) unsigned char s_v[2][kKTile * 128];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * kQTile;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int d = n_heads * kHd;
  const long long ld = 3ll * d;
  const __half* base = qkv + (long long)b * T * ld;
  const __half* gq = base + h * kHd;
  const __half* gk = base + d + h * kHd;
  const __half* gv = base + 2 * d + h * kHd;
  const int n_kt = ceil_div(T, kKTile);

  load_tile_async(s_q, gq, ld, q0, T, tid);
  load_tile_async(s_k[0], gk, ld, 0, T, tid);
  load_tile_async(s_v[0], gv, ld, 0, T, tid);
  cp_async_commit();

  // per-thread state: rows g = lane/4 and g+8 of this warp's 16 query rows
  float o_acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  uint32_t q_frag[4][4];
  const float kLog2e = 1.4426950408889634f;
  const int row_a = q0 + warp * 16 + (lane >> 2);  // query index of accumulator rows 0/1
  const int row_b = row_a + 8;
  float gate_a = 0.f, gate_b = 0.f;
  const float* pb_a = nullptr;
  const float* pb_b = nullptr;
  if (pos_bias) {
    const int ra = min(row_a, T - 1), rb = min(row_b, T - 1);
    gate_a = __ldg(gate + ((long long)b * n_heads + h) * T + ra);
    gate_b = __ldg(gate + ((long long)b * n_heads + h) * T + rb);
    pb_a = pos_bias + ((long long)h * T + ra) * pb_ld;
    pb_b = pos_bias + ((long long)h * T + rb) * pb_ld;
  }

  for (int kt = 0; kt < n_kt; ++kt) {
    const int st = kt & 1;
    cp_async_wait<0>();
    __syncthreads();
    if (kt + 1 < n_kt) {
      load_tile_async(s_k[st ^ 1], gk, ld, (kt + 1) * kKTile, T, tid);
      load_tile_async(s_v[st ^ 1], gv, ld, (kt + 1) * kKTile, T, tid);
      cp_async_commit();
    }
    if (kt == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int r = warp * 16 + (lane & 15);
        const int c = ks * 2 + (lane >> 4);
        ldmatrix_x4(q_frag[ks], smem_u32(s_q + tile_off(r, c)));
      }
    }
    // ---- S = Q K^T (16 x 64 per warp) ----
    float s_acc[8][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      s_acc[nb][0] = s_acc[nb][1] = s_acc[nb][2] = s_acc[nb][3] = 0.f;
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {  // two k-steps per ldmatrix.x4
        uint32_t kf[4];
        const int r = nb * 8 + (lane & 7);
        const int c = kp * 4 + (lane >> 3);
        ldmatrix_x4(kf, smem_u32(s_k[st] + tile_off(r, c)));
        mma_f16_16816(s_acc[nb], q_frag[kp * 2], kf[0], kf[1]);
        mma_f16_16816(s_acc[nb], q_frag[kp * 2 + 1], kf[2], kf[3]);
      }
    }
    // ---- bias, key mask, online softmax ----
    const int key0 = kt * kKTile + (lane & 3) * 2;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = key0 + nb * 8 + (e & 1);
        float v = s_acc[nb][e];
        if (key < T) {
          if (pos_bias) v += (e < 2 ? gate_a * __ldg(pb_a + key) : gate_b * __ldg(pb_b + key));
          v *= kLog2e;
        } else {
          v = -INFINITY;
        }
        s_acc[nb][e] = v;
        mx[e >> 1] = fmaxf(mx[e >> 1], v);
      }
    }
    float scale[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 1));
      mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 2));
      const float m_new = fmaxf(m_run[i], mx[i]);
      scale[i] = (m_run[i] == -INFINITY) ? 0.f : fast_exp2(m_run[i] - m_new);
      m_run[i] = m_new;
      l_run[i] *= scale[i];
    }
    float rs[2] = {0.f, 0.f};
    uint32_t p_frag[4][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      const float p0 = fast_exp2(s_acc[nb][0] - m_run[0]);
      const float p1 = fast_exp2(s_acc[nb][1] - m_run[0]);
      const float p2 = fast_exp2(s_acc[nb][2] - m_run[1]);
      const float p3 = fast_exp2(s_acc[nb][3] - m_run[1]);
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      p_frag[nb >> 1][(nb & 1) * 2 + 0] = pack_f16x2(p0, p1);
      p_frag[nb >> 1][(nb & 1) * 2 + 1] = pack_f16x2(p2, p3);
    }
    l_run[0] += rs[0];
    l_run[1] += rs[1];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      o_acc[nb][0] *= scale[0]; o_acc[nb][1] *= scale[0];
      o_acc[nb][2] *= scale[1]; o_acc[nb][3] *= scale[1];
    }
    // ---- O += P V ----
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {      // 16 keys per step
#pragma unroll
      for (int np = 0; np < 4; ++np) {    // two 8-wide d blocks per ldmatrix.x4.trans
        uint32_t vf[4];
        const int r = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int c = np * 2 + (lane >> 4);
        ldmatrix_x4_trans(vf, smem_u32(s_v[st] + tile_off(r, c)));
        mma_f16_16816(o_acc[np * 2], p_frag[ks], vf[0], vf[1]);
        mma_f16_16816(o_acc[np * 2 + 1], p_frag[ks], vf[2], vf[3]);
      }
    }
  }

  // ---- normalise and store ----
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    l_run[i] += __shfl_xor_sync(0xffffffffu, l_run[i], 1);
    l_run[i] += __shfl_xor_sync(0xffffffffu, l_run[i], 2);
  }
  const float inv_a = 1.0f / l_run[0], inv_b = 1.0f / l_run[1];
  __half* o_base = out + (long long)b * T * d + h * kHd + (lane & 3) * 2;
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) {
    if (row_a < n_query)
      *reinterpret_cast<uint32_t*>(o_base + (long long)row_a * d + nb * 8) =
          pack_f16x2(o_acc[nb][0] * inv_a, o_acc[nb][1] * inv_a);
    if (row_b < n_query)
      *reinterpret_cast<uint32_t*>(o_base + (long long)row_b * d + nb * 8) =
          pack_f16x2(o_acc[nb][2] * inv_b, o_acc[nb][3] * inv_b);
  }
}

int launch_attention_tc5(const void* qkv, int n_windows, int T, int n_heads, int n_query, const float* gate,
                         const float* pos_bias, int pb_ld, int bias_mode, void* out, cudaStream_t st);

}  // namespace segma

using namespace segma;

extern "C" {

int segma_attention(const void* qkv, int n_windows, int T, int n_heads, int n_query, const float* gate,
                    const float* pos_bias, int pos_bias_ld, void* out, void* stream) {
  SEGMA_REQUIRE(n_windows >= 0 && T > 0 && n_heads > 0 && n_query >= 0 && n_query <= T, "segma_attention: bad shape");
  if (n_windows == 0 || n_query == 0) return SEGMA_OK;
  SEGMA_REQUIRE(qkv && out, "segma_attention: NULL buffer");
  SEGMA_REQUIRE((gate == nullptr) == (pos_bias == nullptr), "segma_attention: gate and pos_bias go together");
  SEGMA_REQUIRE(n_heads <= 65535 && n_windows <= 65535, "segma_attention: grid too large");
  SEGMA_REQUIRE(pos_bias == nullptr || pos_bias_ld >= T, "segma_attention: pos_bias_ld %d < T %d", pos_bias_ld, T);
  // tcgen05 path; SEGMA_ATTN_LEGACY=1 (or a bias table whose rows are not 16-byte aligned) selects the
  // warp-level mma.sync kernel below
  static const bool legacy = getenv("SEGMA_ATTN_LEGACY") != nullptr;
  const bool bias_ok = pos_bias == nullptr || (pos_bias_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(pos_bias) & 15) == 0);
  if (!legacy && bias_ok && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0)
    return launch_attention_tc5(qkv, n_windows, T, n_heads, n_query, gate, pos_bias, pos_bias_ld, pos_bias ? 1 : 0, out,
                                (cudaStream_t)stream);
  dim3 grid(ceil_div(n_query, kQTile), n_heads, n_windows);
  attention_kernel<<<grid, kAttnThreads, 0, (cudaStream_t)stream>>>(
      static_cast<const __half*>(qkv), T, n_heads, n_query, gate, pos_bias, pos_bias_ld, static_cast<__half*>(out));
  return launch_status("attention_kernel");
}

int segma_attention_rel(const void* qkv, int n_windows, int T, int n_heads, int n_query, const float* gate,
                        const float* rel_bias, void* out, void* stream) {
  SEGMA_REQUIRE(n_windows >= 0 && T > 0 && n_heads > 0 && n_query >= 0 && n_query <= T, "segma_attention_rel: bad shape");
  if (n_windows == 0 || n_query == 0) return SEGMA_OK;
  SEGMA_REQUIRE(qkv && out && gate && rel_bias, "segma_attention_rel: NULL buffer");
  SEGMA_REQUIRE(n_heads <= 65535 && n_windows <= 65535, "segma_attention_rel: grid too large");
  SEGMA_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0, "segma_attention_rel: qkv must be 16-byte aligned");
  return launch_attention_tc5(qkv, n_windows, T, n_heads, n_query, gate, rel_bias, 0, 2, out, (cudaStream_t)stream);
}

}  // extern "C"
