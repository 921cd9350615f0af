// Fused multi-head self-attention on tcgen05 tensor cores: softmax(Q K^T) V, head_dim 64, fp16 operands,
// fp32 scores / softmax statistics / output accumulation in TMEM.
//
// Replaces torch SDPA as dispatched by WhisperAttention (site-packages/transformers/models/whisper/
// modeling_whisper.py:338-352; 1500 positions, no mask) and torchaudio SelfAttention
// (site-packages/torchaudio/models/wav2vec2/components.py:305-307; 199 positions).
//
// One CTA owns 128 query rows of one (window, head) and walks the keys in tiles of 64; four CTAs are
// resident per SM (48 KB of shared memory, 128 TMEM columns each) so that one CTA's softmax (MUFU-bound)
// overlaps the others' MMAs and barrier hand-offs.  Per key tile:
//   warp 0      TMA: Q once, then the K and V tiles (64 x 64, 128B swizzle; single buffers refilled as soon as the
//               MMA that read them retires), straight from the
//               fused QKV activation through 3-D tensor maps (column block selects q / k / v and the head)
//   warp 1      one lane issues  O += P_{j-1} V_{j-1}  (M128 N64 K64, V as an MN-major B operand) and
//               S = Q K_j^T (M128 N64 K64) back to back, then one tcgen05.commit
//   warps 2-5   one thread per query row: a single tcgen05.ld of its 64 scores, exp2 against the running
//               reference max (the tile is redone, and O/l rescaled, only when a score exceeds it by more than
//               2^8), P -> fp16 into the swizzled A-operand tile
// Scores, probabilities and the output accumulator never touch HBM.
#include "common.cuh"

namespace segma {

constexpr int kAtQ = 128;     // queries per CTA
constexpr int kAtK = 64;      // keys per tile
constexpr int kAtD = 64;      // head dim
constexpr int kAtThreads = 192;
constexpr int kQBytes = kAtQ * 128;   // 128 rows x 64 fp16
constexpr int kKVBytes = kAtK * 128;  // 64 rows x 64 fp16
// Q + K + V + P + barriers = 49 280 B: four CTAs per SM (their 4 x 128 TMEM columns fill the SM's 512).  K and V
// are single-buffered with their own barriers: K_{j+1} is fetched as soon as S_j = Q K_j^T has retired and
// V_{j+1} as soon as O += P_j V_j has, both behind the softmax of the tile in flight.  No alignment slack: the
// dynamic shared-memory window of a kernel without static shared memory starts 1024-byte aligned (checked).
constexpr int kAtSmem = kQBytes + 2 * kKVBytes + kQBytes + 128;
constexpr uint32_t kTmemColsAttn = 128;    // S: columns [0, 64), O: columns [64, 128)
constexpr float kRescaleThreshold = 8.0f;  // log2 units
constexpr uint32_t kWaitHintNs = 2000;     // suspend hint of the mbarrier waits (a completed phase wakes the thread)

// kBias: scores get the WavLM gated relative-position term gate[b,h,i] * pos_bias[h,i,j] added before the softmax
template <bool kBias>
__global__ void __launch_bounds__(kAtThreads, 4)
attention_tc5_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv, int T,
                     int n_heads, int n_query, const float* __restrict__ gate, const float* __restrict__ pos_bias,
                     int pb_ld, __half* __restrict__ out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // 128B-swizzled operand tiles need 1024-byte alignment
  unsigned char* s_q = smem;
  unsigned char* s_k = smem + kQBytes;
  unsigned char* s_v = smem + kQBytes + kKVBytes;
  unsigned char* s_p = smem + kQBytes + 2 * kKVBytes;  // 128 rows x 64 keys
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_p + kQBytes);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* v_full = bars + 2;
  uint64_t* k_empty = bars + 3;
  uint64_t* v_empty = bars + 4;
  uint64_t* mma_done = bars + 5;
  uint64_t* p_ready = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int q0 = blockIdx.x * kAtQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int d = n_heads * kAtD;
  const int n_kt = ceil_div(T, kAtK);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_kv);
    mbar_init(q_full, 1);
    mbar_init(k_full, 1);
    mbar_init(v_full, 1);
    mbar_init(k_empty, 1);
    mbar_init(v_empty, 1);
    mbar_init(mma_done, 1);
    mbar_init(p_ready, 4);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemColsAttn);
    tmem_relinquish();
  }
  tc5_fence_before();
  __syncthreads();
  tc5_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;
  const uint32_t tmem_o = tmem_base + kAtK;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kQBytes);
      tma_load_3d(s_q, &map_q, q_full, h * kAtD, q0, b);
      for (int j = 0; j < n_kt; ++j) {
        mbar_wait_suspend(k_empty, (j & 1) ^ 1, kWaitHintNs);
        mbar_arrive_expect_tx(k_full, kKVBytes);
        tma_load_3d(s_k, &map_kv, k_full, d + h * kAtD, j * kAtK, b);
        mbar_wait_suspend(v_empty, (j & 1) ^ 1, kWaitHintNs);
        mbar_arrive_expect_tx(v_full, kKVBytes);
        tma_load_3d(s_v, &map_kv, v_full, 2 * d + h * kAtD, j * kAtK, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_f16(kAtQ, kAtK, 0, 0);  // S = Q K^T, both K-major
      constexpr uint32_t idesc_o = umma_idesc_f16(kAtQ, kAtD, 0, 1);  // O += P V, V is MN-major
      const uint32_t q_addr = smem_u32(s_q);
      const uint32_t p_addr = smem_u32(s_p);
      mbar_wait_suspend(q_full, 0, kWaitHintNs);
      for (int j = 0; j <= n_kt; ++j) {
        if (j > 0) {
          mbar_wait_suspend(v_full, (j - 1) & 1, kWaitHintNs);
          mbar_wait_suspend(p_ready, (j - 1) & 1, kWaitHintNs);
          tc5_fence_after();
          const uint32_t v_addr = smem_u32(s_v);
#pragma unroll
          for (int ks = 0; ks < kAtK / 16; ++ks) {
            const uint64_t da = umma_desc_k_sw128(p_addr + ks * 32);
            const uint64_t db = umma_desc_mn_sw128(v_addr + ks * 2048, kKVBytes);
            tc5_mma_f16(tmem_o, da, db, idesc_o, (j > 1 || ks > 0) ? 1u : 0u);
          }
          tc5_commit(v_empty);
        }
        if (j < n_kt) {
          mbar_wait_suspend(k_full, j & 1, kWaitHintNs);
          tc5_fence_after();
          const uint32_t k_addr = smem_u32(s_k);
#pragma unroll
          for (int ks = 0; ks < kAtD / 16; ++ks) {
            const uint64_t da = umma_desc_k_sw128(q_addr + ks * 32);
            const uint64_t db = umma_desc_k_sw128(k_addr + ks * 32);
            tc5_mma_f16(tmem_s, da, db, idesc_s, ks > 0 ? 1u : 0u);
          }
          tc5_commit(k_empty);  // K_j is free once S_j has retired
        }
        tc5_commit(mma_done);
      }
    }
  } else {
    // ===================== softmax / correction / epilogue: one thread per query row =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;           // row inside the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const float kLog2e = 1.4426950408889634f;
    float m_run = 0.f;  // reference max, log2 units (set from the first tile)
    float l_run = 0.f;
    // gated relative-position bias of this query row (log2 units): g2 * pb[key]
    float g2 = 0.f;
    const float* pb = nullptr;
    if (kBias) {
      const int qr = min(q0 + row, T - 1);
      g2 = __ldg(gate + ((long long)b * n_heads + h) * T + qr) * kLog2e;
      pb = pos_bias + ((long long)h * T + qr) * pb_ld;
    }
    unsigned char* p_row = s_p + row * 128;
    const int sw = row & 7;
    for (int j = 0; j < n_kt; ++j) {
      mbar_wait_suspend(mma_done, j & 1, kWaitHintNs);
      tc5_fence_after();
      const int n_valid = min(kAtK, T - j * kAtK);  // keys of this tile that exist
      if (j == 0) {  // the first tile fixes the reference max
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t sv[32];
          tmem_ld_32x32(tmem_s + lane_addr + c * 32, sv);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < n_valid) {
              float y = __uint_as_float(sv[i]) * kLog2e;
              if (kBias) y = fmaf(__ldg(pb + c * 32 + i), g2, y);
              mx = fmaxf(mx, y);
            }
        }
        m_run = mx;
      }
      float psum;
      while (true) {
        // running max of the fp16 probabilities (packed pairs): > 2^8 (or inf) means a score exceeded the
        // reference max by more than the threshold
        __half2 pmax2 = __floats2half2_rn(0.f, 0.f);
        psum = 0.f;
        if (n_valid == kAtK) {  // full tile (block-uniform): no key masking
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t sv[32];
            tmem_ld_32x32(tmem_s + lane_addr + c * 32, sv);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint32_t pk[4];
#pragma unroll
              float4 pb0 = make_float4(0.f, 0.f, 0.f, 0.f), pb1 = pb0;
              if (kBias) {  // 8 consecutive keys of this row's bias (rows are padded to a multiple of 4 floats)
                const float4* src = reinterpret_cast<const float4*>(pb + j * kAtK + c * 32 + q * 8);
                pb0 = __ldg(src);
                pb1 = __ldg(src + 1);
              }
              const float off[8] = {pb0.x, pb0.y, pb0.z, pb0.w, pb1.x, pb1.y, pb1.z, pb1.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float o0 = kBias ? fmaf(off[2 * e], g2, -m_run) : -m_run;
                const float o1 = kBias ? fmaf(off[2 * e + 1], g2, -m_run) : -m_run;
                const float p0 = ex2_approx(fmaf(__uint_as_float(sv[q * 8 + 2 * e]), kLog2e, o0));
                const float p1 = ex2_approx(fmaf(__uint_as_float(sv[q * 8 + 2 * e + 1]), kLog2e, o1));
                psum += p0 + p1;
                const __half2 h2 = __floats2half2_rn(p0, p1);
                pmax2 = __hmax2(pmax2, h2);
                pk[e] = *reinterpret_cast<const uint32_t*>(&h2);
              }
              *reinterpret_cast<uint4*>(p_row + (((c * 4 + q) ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          }
        } else {
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t sv[32];
            tmem_ld_32x32(tmem_s + lane_addr + c * 32, sv);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint32_t pk[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int i = c * 32 + q * 8 + 2 * e;
                float o0 = -m_run, o1 = -m_run;
                if (kBias) {
                  if (i < n_valid) o0 = fmaf(__ldg(pb + j * kAtK + i), g2, -m_run);
                  if (i + 1 < n_valid) o1 = fmaf(__ldg(pb + j * kAtK + i + 1), g2, -m_run);
                }
                float p0 = ex2_approx(fmaf(__uint_as_float(sv[q * 8 + 2 * e]), kLog2e, o0));
                float p1 = ex2_approx(fmaf(__uint_as_float(sv[q * 8 + 2 * e + 1]), kLog2e, o1));
                if (i >= n_valid) p0 = 0.f;
                if (i + 1 >= n_valid) p1 = 0.f;
                psum += p0 + p1;
                const __half2 h2 = __floats2half2_rn(p0, p1);
                pmax2 = __hmax2(pmax2, h2);
                pk[e] = *reinterpret_cast<const uint32_t*>(&h2);
              }
              *reinterpret_cast<uint4*>(p_row + (((c * 4 + q) ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          }
        }
        const float pmax = fmaxf(__low2float(pmax2), __high2float(pmax2));
        // tcgen05.ld/st are warp-collective: the rescale decision is taken per warp, each lane with its own factor
        if (!__any_sync(0xffffffffu, pmax > 256.0f)) break;   // 256 = 2^kRescaleThreshold
        // how far this row's scores exceed the reference (log2 units): log2 of its largest probability,
        // recomputed exactly from the scores since the fp16 probability may have overflowed
        float ymax = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t sv[32];
          tmem_ld_32x32(tmem_s + lane_addr + c * 32, sv);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < n_valid) {
              float y = fmaf(__uint_as_float(sv[i]), kLog2e, -m_run);
              if (kBias) y = fmaf(__ldg(pb + j * kAtK + c * 32 + i), g2, y);
              ymax = fmaxf(ymax, y);
            }
        }
        const float grow = fmaxf(ymax, 0.f);
        const float alpha = ex2_approx(-grow);
        m_run += grow;
        l_run *= alpha;
        if (j > 0) {  // O holds contributions of earlier tiles: rescale it in place, then redo this tile
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t ov[32];
            tmem_ld_32x32(tmem_o + lane_addr + c * 32, ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
            tmem_st_32x32(tmem_o + lane_addr + c * 32, ov);
          }
          tmem_st_wait();
        }
      }
      l_run += psum;
      fence_proxy_async_smem();
      tc5_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
    }
    // epilogue: wait for the last P V, normalise, store
    mbar_wait(mma_done, n_kt & 1);
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

int launch_attention_tc5(const void* qkv, int n_windows, int T, int n_heads, int n_query, const float* gate,
                         const float* pos_bias, int pb_ld, void* out, cudaStream_t st) {
  const int d = n_heads * kAtD;
  CUtensorMap map_q, map_kv;
  uint64_t dims[3] = {(uint64_t)3 * d, (uint64_t)T, (uint64_t)n_windows};
  uint64_t strides[3] = {1, (uint64_t)3 * d, (uint64_t)3 * d * T};
  int rc = make_f16_map(&map_q, qkv, 3, dims, strides, kAtQ);
  if (rc != SEGMA_OK) return rc;
  rc = make_f16_map(&map_kv, qkv, 3, dims, strides, kAtK);
  if (rc != SEGMA_OK) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    SEGMA_CUDA_OK(cudaFuncSetAttribute(attention_tc5_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmem));
    SEGMA_CUDA_OK(cudaFuncSetAttribute(attention_tc5_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmem));
    attr_set = true;
  }
  dim3 grid(ceil_div(n_query, kAtQ), n_heads, n_windows);
  if (pos_bias)
    attention_tc5_kernel<true><<<grid, kAtThreads, kAtSmem, st>>>(map_q, map_kv, T, n_heads, n_query, gate, pos_bias,
                                                                 pb_ld, static_cast<__half*>(out));
  else
    attention_tc5_kernel<false><<<grid, kAtThreads, kAtSmem, st>>>(map_q, map_kv, T, n_heads, n_query, nullptr, nullptr,
                                                                  0, static_cast<__half*>(out));
  return launch_status("attention_tc5_kernel");
}

}  // namespace segma
