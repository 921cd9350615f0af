// Fused multi-head self-attention on tcgen05 tensor cores: softmax(Q K^T) V, head_dim 64, fp16 operands,
// fp32 scores / softmax statistics / output accumulation in TMEM.
//
// Replaces torch SDPA as dispatched by WhisperAttention (site-packages/transformers/models/whisper/
// modeling_whisper.py:338-352; 1500 positions, no mask) and torchaudio SelfAttention
// (site-packages/torchaudio/models/wav2vec2/components.py:305-307; 199 positions).
//
// One CTA owns 128 query rows of one (window, head); four CTAs are resident per SM (48 KB of shared memory, 128 TMEM
// columns each) so that one CTA's softmax overlaps the others' MMAs and barrier hand-offs.  K and V arrive in tiles
// of 64 keys; scores are produced and consumed in chunks of 32 keys, in two TMEM buffers:
//   warp 0      TMA: Q once, then the K and V tiles (64 x 64, 128B swizzle, two buffers each), straight from the
//               fused QKV activation through 3-D tensor maps (column block selects q / k / v and the head)
//   warp 1      converged; one elected lane issues, per chunk t,  O += P_t V_t  (M128 N64 K32; P read from tensor
//               memory, V as an MN-major shared-memory B operand) and  S_{t+2} = Q K_{t+2}^T  (M128 N32 K64) into the
//               buffer P_t just left, so the softmax of chunk t+1 overlaps both
//   warps 2-5   one thread per query row, 32 keys at a time: tcgen05.ld of the scores, exp2 against the running
//               reference max (rescaling O and l only when a chunk's probabilities sum to more than 2^10), P -> fp16
//               pairs stored with tcgen05.st over the score columns just consumed
// Scores and probabilities never leave tensor memory / registers; shared memory only carries Q, K and V.
#include <cstdlib>

#include "common.cuh"

namespace segma {

constexpr int kAtQ = 128;     // queries per CTA
constexpr int kAtK = 64;      // keys per K / V tile (one TMA box)
constexpr int kAtC = 32;      // keys per score chunk (one MMA, one softmax step)
constexpr int kAtD = 64;      // head dim
constexpr int kAtThreads = 192;
constexpr int kQBytes = kAtQ * 128;   // 128 rows x 64 fp16
constexpr int kKVBytes = kAtK * 128;  // 64 rows x 64 fp16
// Q + 2 K + 2 V + barriers = 49 280 B: four CTAs per SM (their 4 x 128 TMEM columns fill the SM's 512).  No
// alignment slack: the dynamic shared-memory window of a kernel without static shared memory starts 1024-byte
// aligned (checked).
constexpr int kAtSmem = kQBytes + 4 * kKVBytes + 128;
// TMEM: two score buffers of 32 columns (even / odd chunks) and the 64 output columns.  The probabilities of a
// chunk (fp16 pairs, 16 columns) overwrite the start of its own score buffer.
constexpr uint32_t kTmemColsAttn = 128;
constexpr uint32_t kWaitHintNs = 2000;     // suspend hint of the mbarrier waits (a completed phase wakes the thread)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleSum = 1024.0f;  // chunk sum of probabilities that triggers a new reference max

// Probabilities of one 32-key chunk of a row: pk = fp16 pairs, psum += their fp32 sum.
//   kBias      0 = none; 1 = scores get the WavLM gated relative-position term g2 * pbk[key] (log2 units), pbk a row
//              of the (H, T, T) table in global memory; 2 = the same with pbk pointing into the CTA's shared-memory
//              slice of the Toeplitz vector (any alignment)
//   kPolyMask  bit (4 q + e) set = key pair e of the q-th group of 8 keys takes its exponentials from ex2_poly_pair
//              (FMA pipe) instead of the MUFU
//   kMasked    only the first nv keys exist (last tile of a row)
template <int kBias, uint32_t kPolyMask, bool kMasked>
__device__ __forceinline__ void softmax_chunk(const uint32_t (&sv)[32], float m_run, float g2, const float* pbk,
                                              int nv, uint32_t (&pk)[16], float& psum) {
  const uint64_t k2 = f2_pack(kLog2e, kLog2e);
  const uint64_t nm2 = f2_pack(-m_run, -m_run);
  uint64_t psum2 = f2_pack(0.f, 0.f);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float off[8];
    if (kBias == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) off[i] = (!kMasked || q * 8 + i < nv) ? pbk[q * 8 + i] : 0.f;
    } else if (kBias == 1) {
      if (!kMasked) {  // rows of the bias table are padded to a multiple of 4 floats
        const float4 a = __ldg(reinterpret_cast<const float4*>(pbk + q * 8));
        const float4 b = __ldg(reinterpret_cast<const float4*>(pbk + q * 8) + 1);
        off[0] = a.x; off[1] = a.y; off[2] = a.z; off[3] = a.w;
        off[4] = b.x; off[5] = b.y; off[6] = b.z; off[7] = b.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) off[i] = (q * 8 + i < nv) ? __ldg(pbk + q * 8 + i) : 0.f;
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = q * 8 + 2 * e;
      const uint64_t off2 = kBias ? f2_pack(fmaf(off[2 * e], g2, -m_run), fmaf(off[2 * e + 1], g2, -m_run)) : nm2;
      const uint64_t x2 = f2_fma(f2_pack(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])), k2, off2);
      float p0, p1;
      if ((kPolyMask >> (4 * q + e)) & 1u) {
        ex2_poly_pair(x2, p0, p1);
      } else {
        float x0, x1;
        f2_unpack(x2, x0, x1);
        p0 = ex2_approx(x0);
        p1 = ex2_approx(x1);
      }
      if (kMasked) {
        if (i >= nv) p0 = 0.f;
        if (i + 1 >= nv) p1 = 0.f;
      }
      psum2 = f2_add(psum2, f2_pack(p0, p1));
      pk[q * 4 + e] = pack_f16x2(p0, p1);
    }
  }
  float s0, s1;
  f2_unpack(psum2, s0, s1);
  psum += s0 + s1;
}

// largest score of the chunk in log2 units relative to `ref`, keys >= nv ignored
template <int kBias>
__device__ __forceinline__ float chunk_max(const uint32_t (&sv)[32], float ref, float g2, const float* pbk, int nv) {
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i < nv) {
      float y = fmaf(__uint_as_float(sv[i]), kLog2e, -ref);
      if (kBias) y = fmaf(pbk[i], g2, y);
      mx = fmaxf(mx, y);
    }
  return mx;
}

template <int kBias, uint32_t kPolyMask>
__global__ void __launch_bounds__(kAtThreads, 4)
attention_tc5_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv, int T,
                     int n_heads, int n_query, const float* __restrict__ gate, const float* __restrict__ pos_bias,
                     int pb_ld, __half* __restrict__ out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // 128B-swizzled operand tiles need 1024-byte alignment
  unsigned char* s_q = smem;
  unsigned char* s_k = smem + kQBytes;                 // two K tiles
  unsigned char* s_v = smem + kQBytes + 2 * kKVBytes;  // two V tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kQBytes + 4 * kKVBytes);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;   // [2] K tile landed
  uint64_t* v_full = bars + 3;   // [2]
  uint64_t* k_empty = bars + 5;  // [2] both score chunks of the tile have retired
  uint64_t* v_empty = bars + 7;  // [2] both P V chunks of the tile have retired
  uint64_t* s_full = bars + 9;   // [2] scores of a chunk are in their buffer
  uint64_t* p_ready = bars + 11; // [2] probabilities of a chunk are in their buffer (4 warps arrive)
  uint64_t* pv_done = bars + 13; // one phase per retired P V chunk
  uint64_t* o_final = bars + 14; // the last P V chunk has retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);
  // kBias == 2: s_rel[x] = rel[h][x + (T - 1) - (q0 + 127)], so that row r and key j read s_rel[j - r + 127]
  float* s_rel = reinterpret_cast<float*>(smem + kAtSmem);

  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int q0 = blockIdx.x * kAtQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int d = n_heads * kAtD;
  const int n_kt = ceil_div(T, kAtK);
  const int n_ch = ceil_div(T, kAtC);  // chunks that hold at least one key

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_kv);
    for (int i = 0; i < 15; ++i) mbar_init(bars + i, (i == 11 || i == 12) ? 4 : 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemColsAttn);
    tmem_relinquish();
  }
  if (kBias == 2) {
    const float* rel = pos_bias + (long long)blockIdx.y * (2 * T - 1);
    for (int x = threadIdx.x; x < T + kAtQ; x += kAtThreads) {
      const int ri = x + (T - 1) - (blockIdx.x * kAtQ + kAtQ - 1);
      s_rel[x] = (ri >= 0 && ri < 2 * T - 1) ? __ldg(rel + ri) : 0.f;
    }
  }
  tc5_fence_before();
  __syncthreads();
  tc5_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;             // two score / probability buffers of 32 columns
  const uint32_t tmem_o = tmem_base + 2 * kAtC;  // 64 output columns

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, kQBytes);
      tma_load_3d(s_q, &map_q, q_full, h * kAtD, q0, b);
    }
    __syncwarp();
#pragma unroll 1
    for (int j = 0; j < n_kt; ++j) {
      const int s = j & 1, ph = (j >> 1) & 1;
      mbar_wait_suspend(k_empty + s, ph ^ 1, kWaitHintNs);
      if (elect_one()) {
        mbar_arrive_expect_tx(k_full + s, kKVBytes);
        tma_load_3d(s_k + s * kKVBytes, &map_kv, k_full + s, d + h * kAtD, j * kAtK, b);
      }
      __syncwarp();
      mbar_wait_suspend(v_empty + s, ph ^ 1, kWaitHintNs);
      if (elect_one()) {
        mbar_arrive_expect_tx(v_full + s, kKVBytes);
        tma_load_3d(s_v + s * kKVBytes, &map_kv, v_full + s, 2 * d + h * kAtD, j * kAtK, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // The whole warp walks the chunk loop converged (all lanes poll the barriers); one elected lane issues.
    constexpr uint32_t idesc_s = umma_idesc_f16(kAtQ, kAtC, 0, 0);  // S = Q K^T, both K-major, 32 keys
    constexpr uint32_t idesc_o = umma_idesc_f16(kAtQ, kAtD, 0, 1);  // O += P V, V is MN-major
    // operand descriptors differ only in the start-address field (bytes >> 4): built once, then advanced by adds
    const uint64_t q_desc = umma_desc_k_sw128(smem_u32(s_q));
    const uint64_t k_desc = umma_desc_k_sw128(smem_u32(s_k));
    const uint64_t v_desc = umma_desc_mn_sw128(smem_u32(s_v), kKVBytes);
    // scores of chunk t into buffer t & 1; u = t mod 4 fixes the TMEM buffer and the K stage at compile time
    auto issue_scores = [&](int t, int u) {
      const int half = u & 1, s = (u >> 1) & 1;
      if (half == 0) mbar_wait_suspend(k_full + s, (t >> 2) & 1, kWaitHintNs);
      tc5_fence_after();
      if (elect_one()) {
        const uint64_t kd = k_desc + static_cast<uint32_t>((s * kKVBytes + half * (kAtC * 128)) >> 4);
#pragma unroll
        for (int ks = 0; ks < kAtD / 16; ++ks)  // 16 head-dim elements = 32 bytes along K
          tc5_mma_f16(tmem_s + half * kAtC, q_desc + 2 * ks, kd + 2 * ks, idesc_s, ks > 0 ? 1u : 0u);
        if (half == 1 || t == n_ch - 1) tc5_commit(k_empty + s);  // the K tile is free once its last chunk retires
        tc5_commit(s_full + half);
      }
      __syncwarp();
    };
    mbar_wait_suspend(q_full, 0, kWaitHintNs);
    issue_scores(0, 0);
    if (n_ch > 1) issue_scores(1, 1);
#pragma unroll 1
    for (int t0 = 0; t0 < n_ch; t0 += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int t = t0 + u;
        if (t >= n_ch) break;
        const int half = u & 1, s = (u >> 1) & 1;
        if (half == 0) mbar_wait_suspend(v_full + s, (t >> 2) & 1, kWaitHintNs);
        mbar_wait_suspend(p_ready + half, (t >> 1) & 1, kWaitHintNs);
        tc5_fence_after();
        if (elect_one()) {
          const uint64_t vd = v_desc + static_cast<uint32_t>((s * kKVBytes + half * (kAtC * 128)) >> 4);
#pragma unroll
          for (int ks = 0; ks < kAtC / 16; ++ks)  // 16 keys = 8 columns of fp16 pairs = 16 V rows of 128 bytes
            tc5_mma_f16_ts(tmem_o, tmem_s + half * kAtC + ks * 8, vd + ks * (2048 >> 4), idesc_o,
                           (t > 0 || ks > 0) ? 1u : 0u);
          if (half == 1 || t == n_ch - 1) tc5_commit(v_empty + s);
          tc5_commit(pv_done);
          if (t == n_ch - 1) tc5_commit(o_final);
        }
        __syncwarp();
        if (t + 2 < n_ch) issue_scores(t + 2, (u + 2) & 3);  // same buffer: ordered behind the P V that reads it
      }
    }
  } else {
    // ===================== softmax / correction / epilogue: one thread per query row =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;           // row inside the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    float m_run = 0.f;  // reference max, log2 units (set from the first chunk)
    float l_run = 0.f;
    // gated relative-position bias of this query row (log2 units): g2 * pb[key]
    float g2 = 0.f;
    const float* pb = nullptr;
    if (kBias) {
      const int qr = min(q0 + row, T - 1);
      g2 = __ldg(gate + ((long long)b * n_heads + h) * T + qr) * kLog2e;
      pb = kBias == 2 ? s_rel + (kAtQ - 1 - row) : pos_bias + ((long long)h * T + qr) * pb_ld;
    }
#pragma unroll 1
    for (int t = 0; t < n_ch; ++t) {
      const int half = t & 1;
      const uint32_t buf = tmem_s + lane_addr + half * kAtC;
      mbar_wait_suspend(s_full + half, (t >> 1) & 1, kWaitHintNs);
      tc5_fence_after();
      const int nv = min(kAtC, T - t * kAtC);  // keys of this chunk that exist (block-uniform)
      const float* pbk = kBias ? pb + t * kAtC : nullptr;
      uint32_t sv[32];
      tmem_ld_32x32(buf, sv);
      tmem_ld_wait();
      if (t == 0) m_run = chunk_max<kBias>(sv, 0.f, g2, pbk, nv);  // the first chunk fixes the reference
      uint32_t pk[16];
      float psum;
      for (int pass = 0;; ++pass) {  // at most one redo: after it every score is at or below the reference
        // The probabilities are non-negative, so their sum bounds each of them: a chunk sum above 2^10 (or inf / NaN)
        // is the only way a score can have exceeded the reference max by enough to threaten the fp16 range.
        psum = 0.f;
        if (nv == kAtC) softmax_chunk<kBias, kPolyMask, false>(sv, m_run, g2, pbk, kAtC, pk, psum);
        else softmax_chunk<kBias, 0u, true>(sv, m_run, g2, pbk, nv, pk, psum);
        // tcgen05.ld/st are warp-collective: the rescale decision is taken per warp, each lane with its own factor
        if (pass == 1 || !__any_sync(0xffffffffu, !(psum <= kRescaleSum))) break;
        // how far this row's scores exceed the reference (log2 units), exactly, from the scores still in registers
        const float grow = fmaxf(chunk_max<kBias>(sv, m_run, g2, pbk, nv), 0.f);
        const float alpha = ex2_approx(-grow);
        m_run += grow;
        l_run *= alpha;
        if (t > 0) {  // O holds earlier chunks: wait until the last of them has retired, rescale in place
          // (chunk t - 2 retired before these scores did, so the barrier is in phase t - 1 or t: parity is unambiguous)
          mbar_wait(pv_done, (t - 1) & 1);
          tc5_fence_after();
#pragma unroll 1
          for (int cc = 0; cc < 2; ++cc) {
            uint32_t ov[32];
            tmem_ld_32x32(tmem_o + lane_addr + cc * 32, ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
            tmem_st_32x32(tmem_o + lane_addr + cc * 32, ov);
          }
          tmem_st_wait();
        }
      }
      l_run += psum;
      tmem_st_32x16(buf, pk);  // over the score columns this thread has consumed
      tmem_st_wait();
      tc5_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready + half);
    }
    // epilogue: wait for the last P V, normalise, store
    mbar_wait(o_final, 0);
    tc5_fence_after();
    const int q_row = q0 + row;
    const float inv_l = 1.0f / l_run;
    __half* o_ptr = out + ((long long)b * T + q_row) * d + h * kAtD;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t ov[32];
      tmem_ld_32x32(tmem_o + lane_addr + c * 32, ov);
      tmem_ld_wait();
      if (q_row < n_query) {
        SEGMA_DEV_ASSERT(q_row < T && h < n_heads && l_run > 0.f);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 pk;
          pk.x = pack_f16x2(__uint_as_float(ov[8 * q]) * inv_l, __uint_as_float(ov[8 * q + 1]) * inv_l);
          pk.y = pack_f16x2(__uint_as_float(ov[8 * q + 2]) * inv_l, __uint_as_float(ov[8 * q + 3]) * inv_l);
          pk.z = pack_f16x2(__uint_as_float(ov[8 * q + 4]) * inv_l, __uint_as_float(ov[8 * q + 5]) * inv_l);
          pk.w = pack_f16x2(__uint_as_float(ov[8 * q + 6]) * inv_l, __uint_as_float(ov[8 * q + 7]) * inv_l);
          reinterpret_cast<uint4*>(o_ptr + c * 32)[q] = pk;
        }
      }
    }
  }

  tc5_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc5_fence_after();
    tmem_dealloc(tmem_base, kTmemColsAttn);
  }
}

int make_f16_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                 int box_rows);

// bias_mode: 0 none, 1 pos_bias = (H, T, T) table with row stride pb_ld, 2 pos_bias = (H, 2T - 1) Toeplitz vectors
int launch_attention_tc5(const void* qkv, int n_windows, int T, int n_heads, int n_query, const float* gate,
                         const float* pos_bias, int pb_ld, int bias_mode, void* out, cudaStream_t st) {
  const int d = n_heads * kAtD;
  CUtensorMap map_q, map_kv;
  uint64_t dims[3] = {(uint64_t)3 * d, (uint64_t)T, (uint64_t)n_windows};
  uint64_t strides[3] = {1, (uint64_t)3 * d, (uint64_t)3 * d * T};
  int rc = make_f16_map(&map_q, qkv, 3, dims, strides, kAtQ);
  if (rc != SEGMA_OK) return rc;
  rc = make_f16_map(&map_kv, qkv, 3, dims, strides, kAtK);
  if (rc != SEGMA_OK) return rc;
  static PerDeviceFlag attr_set;
  // SEGMA_ATTN_POLY=1 moves a quarter of the exponentials from the MUFU to the FMA pipe (ex2_poly_pair).  Measured
  // slower on B200 (0.370 vs 0.333 ms for 32 windows): the softmax warps are latency-bound, not MUFU-bound, so the
  // longer instruction stream costs more than the freed MUFU slots give back.  Kept as a switch for re-measurement.
  static unsigned poly = 0;
  constexpr int kRelMaxT = 1024;  // the Toeplitz slice (T + 128 floats) must leave room for four CTAs per SM
  if (!attr_set.here()) {
    SEGMA_CUDA_OK(cudaFuncSetAttribute(attention_tc5_kernel<0, 0x0000u>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmem));
    SEGMA_CUDA_OK(cudaFuncSetAttribute(attention_tc5_kernel<0, 0x2222u>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmem));
    SEGMA_CUDA_OK(cudaFuncSetAttribute(attention_tc5_kernel<0, 0x2A2Au>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmem));
    SEGMA_CUDA_OK(cudaFuncSetAttribute(attention_tc5_kernel<0, 0xAAAAu>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmem));
    SEGMA_CUDA_OK(cudaFuncSetAttribute(attention_tc5_kernel<1, 0u>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmem));
    SEGMA_CUDA_OK(cudaFuncSetAttribute(attention_tc5_kernel<2, 0u>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kAtSmem + 4 * (kRelMaxT + kAtQ)));
    const char* env = getenv("SEGMA_ATTN_POLY");
    poly = env ? (unsigned)strtoul(env, nullptr, 0) : 0u;
    attr_set.here() = true;
  }
  if (bias_mode == 2 && T > kRelMaxT) {
    set_last_error("segma_attention_rel: T=%d exceeds %d", T, kRelMaxT);
    return SEGMA_ERR_UNSUPPORTED;
  }
  dim3 grid(ceil_div(n_query, kAtQ), n_heads, n_windows);
#define SEGMA_ATTN_LAUNCH(BIAS, MASK, SMEM)                                                                      \
  attention_tc5_kernel<BIAS, MASK><<<grid, kAtThreads, SMEM, st>>>(map_q, map_kv, T, n_heads, n_query, gate,     \
                                                                    pos_bias, pb_ld, static_cast<__half*>(out))
  if (bias_mode == 2) SEGMA_ATTN_LAUNCH(2, 0u, kAtSmem + 4 * (T + kAtQ));
  else if (bias_mode == 1) SEGMA_ATTN_LAUNCH(1, 0u, kAtSmem);
  else if (poly == 1 || poly == 0x2222u) SEGMA_ATTN_LAUNCH(0, 0x2222u, kAtSmem);  // 4 of 16 key pairs
  else if (poly == 0x2A2Au) SEGMA_ATTN_LAUNCH(0, 0x2A2Au, kAtSmem);                // 6 of 16
  else if (poly == 0xAAAAu) SEGMA_ATTN_LAUNCH(0, 0xAAAAu, kAtSmem);                // 8 of 16
  else SEGMA_ATTN_LAUNCH(0, 0x0000u, kAtSmem);
#undef SEGMA_ATTN_LAUNCH
  return launch_status("attention_tc5_kernel");
}

}  // namespace segma
