// fp16 GEMM / implicit-GEMM Conv1d on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// C[M, N] = epilogue(A[M, K] * W[N, K]^T) with fp32 accumulation.  Replaces torch's dispatch of
// nn.Linear (q/k/v/out projections, fc1/fc2; site-packages/transformers/models/whisper/modeling_whisper.py:
// 279-282,310,331-332,355,376-377 and site-packages/torchaudio/models/wav2vec2/components.py:265-268,324-326)
// and nn.Conv1d of the Whisper stem (modeling_whisper.py:619-625) / wav2vec2 feature extractor
// (components.py:77-99) to cuBLASLt / cuDNN.
//
// Structure (persistent, warp-specialised, one CTA per SM; 256-wide tiles run on CTA pairs, cta_group::2, M = 256):
//   warp 0      TMA producer: 128B-swizzled A (128 x 64) and W (BN x 64, or half of it per CTA of a pair) tiles into a
//               multi-stage ring
//   warp 1      allocates TMEM, one lane issues tcgen05.mma (M=128 or 256, N=BN, K=16) per 32 bytes of K
//   warps 2-..  8 or 16 epilogue warps: tcgen05.ld of the fp32 accumulator (double-buffered in TMEM so the next
//               tile's MMAs overlap), + bias, exact-erf GELU, + fp32 residual / position table, fp16 or fp32 store
//               through a per-warp shared-memory transpose (fp16 outputs without a residual are narrowed first)
// A strided Conv1d is the same loop with the K axis split into taps: tap j of a stride-s convolution reads
// the activation map (rows merged s at a time) at column block (j % s) * C and row offset j / s, so no
// im2col buffer is ever written.
#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace segma {

constexpr int kBM = 128;
constexpr int kBK = 64;  // 128 bytes of fp16: one swizzle row
// epilogue warps: kEW / 4 per TMEM lane quadrant, interleaving 32-column chunks (8, or 16 for the ALU-heavy
// GELU epilogue without residual, whose smaller register footprint leaves room for the extra warps)
constexpr int kEpiPitch = 36;  // floats per staged accumulator row (32 + 4 pad: conflict-free 128-bit access)
constexpr int kEpiRows = 16;   // rows staged per round (half a warp's accumulator rows)
constexpr int kEpiF16Pitch = 80;  // bytes per staged fp16 row (32 values + 16 B pad: conflict-free 128-bit writes)
constexpr int kEpiWarpBytes = 32 * kEpiF16Pitch;  // per-warp staging tile: 32 fp16 rows (>= 16 fp32 rows of 144 B)
static_assert(kEpiWarpBytes >= kEpiRows * kEpiPitch * 4, "staging tile too small");
constexpr int kFlagEpiRegs = 1 << 16;  // internal: bias / GELU on the accumulator registers, fp16 staging

struct GemmKernelArgs {
  int batch, rows_per_batch, tiles_per_batch;
  int n, n_tiles;
  // K axis: taps * kb_per_tap blocks of 64
  int taps, kb_per_tap, conv_stride;
  int a_tap_elems;   // channels per tap in the A map (column offset of tap j = (j % s) * a_tap_elems)
  int w_tap_elems;   // K extent of one tap in the W map
  int a_ntile_off;   // extra A column offset per N tile (grouped convolution), else 0
  const float* bias;
  const float* add_src;
  long long add_batch_rows;
  void* out;
  long long out_batch_rows, out_row_offset, ldo;
  int flags;
};

// kCta2: the tile is computed by a CTA pair (cta_group::2): M = 256 (this CTA's 128 rows plus the peer's), each CTA
// stages only its half of the W tile, and the tensor cores of both SMs read both halves -- 2/3 of the shared-memory
// operand traffic of the single-CTA tile, and two more pipeline stages in the same shared memory.
template <int BN, bool kCta2 = false, int kEW = 8>
struct GemmCfg {
  static constexpr int kThreads = 64 + 32 * kEW;
  static constexpr int kStages = kCta2 ? (kEW > 8 ? 5 : 6) : (BN == 256 ? 4 : (BN == 192 ? 5 : 6));
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBRows = kCta2 ? BN / 2 : BN;  // W rows staged by this CTA
  static constexpr int kBBytes = kBRows * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = BN == 128 ? 256 : 512;
  static constexpr int kAccStride = BN == 192 ? 256 : BN;  // column offset between the two accumulators
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ +
                                    kEW * kEpiWarpBytes /*epilogue transpose tiles*/;
};

template <int BN, bool kCta2, int kEW, bool kAddSrc>
__global__ void __launch_bounds__(64 + 32 * kEW, 1)
gemm_tc5_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                const GemmKernelArgs p) {
  using Cfg = GemmCfg<BN, kCta2, kEW>;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + ((1024 - (raw_addr & 1023)) & 1023);
  unsigned char* smem_a = smem;
  unsigned char* smem_b = smem + Cfg::kStages * Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* tmem_full = bars + 2 * Cfg::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = lane_id();
  const int num_kb = p.taps * p.kb_per_tap;
  // work units: single tiles, or (kCta2) pairs of consecutive M tiles of one batch sharing an N tile
  const int rank = kCta2 ? static_cast<int>(cluster_ctarank()) : 0;
  const int units_per_batch = kCta2 ? (p.tiles_per_batch + 1) / 2 : p.tiles_per_batch;
  const int num_tiles = p.batch * units_per_batch * p.n_tiles;
  const int unit0 = kCta2 ? (blockIdx.x >> 1) : blockIdx.x;
  const int unit_step = kCta2 ? (gridDim.x >> 1) : gridDim.x;
  // unit -> (n_tile, batch, first row of this CTA's 128-row M tile)
  auto locate = [&](int unit, int& n_tile, int& b, int& r0) {
    n_tile = unit % p.n_tiles;
    const int mu = unit / p.n_tiles;
    b = mu / units_per_batch;
    const int m_in_batch = mu - b * units_per_batch;
    r0 = (kCta2 ? 2 * m_in_batch + rank : m_in_batch) * kBM;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_w);
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(full_bar + i, 1);
      mbar_init(empty_bar + i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tmem_full + i, 1);
      mbar_init(tmem_empty + i, kCta2 ? 2 * kEW : kEW);  // the leader collects both CTAs' epilogues
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    if (kCta2) {
      tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish();
    }
  }
  tc5_fence_before();
  if (kCta2) cluster_sync_all(); else __syncthreads();  // the peer's barriers and TMEM must exist before any traffic
  tc5_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = unit0; tile < num_tiles; tile += unit_step) {
        int n_tile, b, r0;
        locate(tile, n_tile, b, r0);
        const int n0 = n_tile * BN + rank * Cfg::kBRows;  // this CTA's rows of the W tile
        for (int tap = 0; tap < p.taps; ++tap) {
          const int a_col = (tap % p.conv_stride) * p.a_tap_elems + n_tile * p.a_ntile_off;
          const int a_row = r0 + tap / p.conv_stride;
          const int w_col = tap * p.w_tap_elems;
          for (int kk = 0; kk < p.kb_per_tap; ++kk) {
            mbar_wait(empty_bar + stage, phase ^ 1);
            if (kCta2) {
              // both CTAs' bytes are credited to the leader's barrier, which is the one the MMA thread waits on
              if (rank == 0) mbar_arrive_expect_tx(full_bar + stage, 2 * Cfg::kStageBytes);
              tma_load_3d_pair(smem_a + stage * Cfg::kABytes, &map_a, full_bar + stage, a_col + kk * kBK, a_row, b);
              tma_load_2d_pair(smem_b + stage * Cfg::kBBytes, &map_w, full_bar + stage, w_col + kk * kBK, n0);
            } else {
              mbar_arrive_expect_tx(full_bar + stage, Cfg::kStageBytes);
              tma_load_3d(smem_a + stage * Cfg::kABytes, &map_a, full_bar + stage, a_col + kk * kBK, a_row, b);
              tma_load_2d(smem_b + stage * Cfg::kBBytes, &map_w, full_bar + stage, w_col + kk * kBK, n0);
            }
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && rank == 0) {  // in a CTA pair only the leader issues
      constexpr uint32_t idesc = umma_idesc_f16(kCta2 ? 2 * kBM : kBM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = unit0; tile < num_tiles; tile += unit_step, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(tmem_empty + as, aphase ^ 1);
        tc5_fence_after();
        const uint32_t d_tmem = tmem_base + as * Cfg::kAccStride;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar + stage, phase);
          tc5_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * Cfg::kABytes);
          const uint32_t b_addr = smem_u32(smem_b + stage * Cfg::kBBytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t da = umma_desc_k_sw128(a_addr + k * 32);
            const uint64_t db = umma_desc_k_sw128(b_addr + k * 32);
            if (kCta2) tc5_mma_f16_pair(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            else tc5_mma_f16(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          // frees the smem stage (in both CTAs of a pair) once these MMAs retire
          if (kCta2) tc5_commit_pair(empty_bar + stage); else tc5_commit(empty_bar + stage);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        if (kCta2) tc5_commit_pair(tmem_full + as); else tc5_commit(tmem_full + as);  // accumulator complete
      }
    }
  } else {
    // ===================== epilogue warps =====================
    // TMEM hands each lane one accumulator row; a 32 x 32 chunk is transposed through a padded
    // shared-memory tile so that global loads/stores run along rows (8 lanes x 16 B per row).
    const int quad = warp & 3;          // TMEM lane quadrant this warp may access
    const int cgrp = (warp - 2) >> 2;   // which of the quadrant's kEW/4 warps: chunks cgrp, cgrp + kEW/4, ...
    const bool out_f32 = (p.flags & SEGMA_GEMM_OUT_F32) != 0;
    const bool do_gelu = (p.flags & SEGMA_GEMM_GELU) != 0;
    unsigned char* stg_bytes = smem + Cfg::kStages * Cfg::kStageBytes + 256 + (warp - 2) * kEpiWarpBytes;
    float* stg = reinterpret_cast<float*>(stg_bytes);
    const bool epi_regs = !kAddSrc && !out_f32 && (p.flags & kFlagEpiRegs) != 0;
    const int sub_r = lane >> 3;   // row within a group of 4
    const int c4 = lane & 7;       // float4 column within the 32-column chunk
    int it = 0;
    for (int tile = unit0; tile < num_tiles; tile += unit_step, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      int n_tile, b, r0;
      locate(tile, n_tile, b, r0);
      const int r_base = r0 + quad * 32;  // first row of this warp
      const int n0 = n_tile * BN;
      const long long out_row0 = (long long)b * p.out_batch_rows + p.out_row_offset + r_base;
      long long src_base = 0;
      if (kAddSrc && p.add_src) src_base = (long long)b * p.add_batch_rows + r_base;
      // up to 16 warps wait here for most of a tile's main loop: suspended polling instead of a hot spin is worth
      // 1.5-3 % of sustained throughput under the board power cap
      mbar_wait_suspend(tmem_full + as, aphase, 4000);
      tc5_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * Cfg::kAccStride;
#pragma unroll 1
      for (int chunk = cgrp; chunk < BN / 32; chunk += kEW / 4) {
        const int nc = n0 + chunk * 32;
        if (nc >= p.n) break;  // warp-uniform
        // issue the residual / position-table loads first so their latency hides behind the TMEM read
        float4 src4[kAddSrc ? 8 : 1];
        if (kAddSrc && p.add_src) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const int rr = g * 4 + sub_r;
            const long long sr = src_base + rr;
            const bool ok = r_base + rr < p.rows_per_batch;
            src4[g] = ok ? *(reinterpret_cast<const float4*>(p.add_src + sr * p.n + nc) + c4)
                         : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        uint32_t acc[32];
        if (!kAddSrc && epi_regs) {
          // fp16 output without a residual: bias and GELU run on the accumulator registers (this lane's row, 32
          // independent columns), the result is packed to fp16 first and only then transposed through shared memory
          // (half the staging bytes of the fp32 route) for row-wise 128-bit stores
          tmem_ld_32x32(t_addr + chunk * 32, acc);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias) bq = __ldg(reinterpret_cast<const float4*>(p.bias + nc) + q);  // same address in every lane
            uint64_t v01 = f2_add(f2_pack(__uint_as_float(acc[4 * q]), __uint_as_float(acc[4 * q + 1])), f2_pack(bq.x, bq.y));
            uint64_t v23 = f2_add(f2_pack(__uint_as_float(acc[4 * q + 2]), __uint_as_float(acc[4 * q + 3])), f2_pack(bq.z, bq.w));
            if (do_gelu) { v01 = gelu_erf_pair(v01); v23 = gelu_erf_pair(v23); }
            float a, b;
            f2_unpack(v01, a, b);
            pk[2 * q] = pack_f16x2(a, b);
            f2_unpack(v23, a, b);
            pk[2 * q + 1] = pack_f16x2(a, b);
          }
          uint4* my_row = reinterpret_cast<uint4*>(stg_bytes + lane * kEpiF16Pitch);
#pragma unroll
          for (int j = 0; j < 4; ++j) my_row[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          __syncwarp();
          const int seg = lane & 3;  // 16-byte segment of the 64-byte row
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int rr = it * 8 + (lane >> 2);
            const uint4 v = *reinterpret_cast<const uint4*>(stg_bytes + rr * kEpiF16Pitch + seg * 16);
            if (r_base + rr < p.rows_per_batch) {
              SEGMA_DEV_ASSERT(nc + seg * 8 + 8 <= p.n && out_row0 + rr >= 0 && b < p.batch);
              *reinterpret_cast<uint4*>(static_cast<__half*>(p.out) + (out_row0 + rr) * p.ldo + nc + seg * 8) = v;
            }
          }
          __syncwarp();
          continue;
        }
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + nc) + c4);
        tmem_ld_32x32(t_addr + chunk * 32, acc);
        tmem_ld_wait();
#pragma unroll
        for (int round = 0; round < 2; ++round) {
          // lanes [16*round, 16*round+16) stage their accumulator rows, then the whole warp stores them row-wise
          if ((lane >> 4) == round) {
            float4* my_row = reinterpret_cast<float4*>(stg + (lane & 15) * kEpiPitch);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              my_row[j] = make_float4(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]),
                                      __uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3]));
          }
          __syncwarp();
#pragma unroll
          for (int gi = 0; gi < 4; ++gi) {
            const int g = round * 4 + gi;
            const int rr = g * 4 + sub_r;
            float4 v = *reinterpret_cast<const float4*>(stg + (gi * 4 + sub_r) * kEpiPitch + c4 * 4);
            {  // bias and GELU on packed fp32 pairs: the epilogue of a GELU GEMM is issue-bound
              uint64_t v01 = f2_add(f2_pack(v.x, v.y), f2_pack(bias4.x, bias4.y));
              uint64_t v23 = f2_add(f2_pack(v.z, v.w), f2_pack(bias4.z, bias4.w));
              if (do_gelu) { v01 = gelu_erf_pair(v01); v23 = gelu_erf_pair(v23); }
              f2_unpack(v01, v.x, v.y);
              f2_unpack(v23, v.z, v.w);
            }
            if (kAddSrc && p.add_src) { v.x += src4[g].x; v.y += src4[g].y; v.z += src4[g].z; v.w += src4[g].w; }
            if (r_base + rr < p.rows_per_batch) {
              SEGMA_DEV_ASSERT(nc + c4 * 4 + 4 <= p.n && out_row0 + rr >= 0 && b < p.batch && n_tile < p.n_tiles);
              const long long o = (out_row0 + rr) * p.ldo + nc;
              if (out_f32) {
                reinterpret_cast<float4*>(static_cast<float*>(p.out) + o)[c4] = v;
              } else {
                uint2 q;
                q.x = pack_f16x2(v.x, v.y);
                q.y = pack_f16x2(v.z, v.w);
                reinterpret_cast<uint2*>(static_cast<__half*>(p.out) + o)[c4] = q;
              }
            }
          }
          __syncwarp();
        }
      }
      tc5_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kCta2) mbar_arrive_leader(tmem_empty + as); else mbar_arrive(tmem_empty + as);
      }
    }
  }

  tc5_fence_before();
  if (kCta2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc5_fence_after();
    if (kCta2) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols); else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---- host side ----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// fp16 tensor map with a (64 x box_rows [x 1]) box and 128B swizzle
int make_f16_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                  int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return SEGMA_ERR_CUDA;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t box[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    estr[i] = 1;
    box[i] = 1;
    if (i > 0) gstr[i - 1] = strides_elems[i] * 2;
  }
  box[0] = kBK;
  box[1] = box_rows;  // (64 x box_rows) tile; higher dims have a box of 1
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu x %llu, stride %llu B)", (int)r,
                   rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                   (unsigned long long)(rank > 1 ? strides_elems[1] * 2 : 0));
    return SEGMA_ERR_CUDA;
  }
  return SEGMA_OK;
}

// CTAs of the persistent grid: one per SM, or fewer when SEGMA_GEMM_MAX_CTAS leaves SMs to a concurrent kernel
static int gemm_cta_limit() {
  static int env_limit = -1;
  if (env_limit < 0) {
    const char* env = getenv("SEGMA_GEMM_MAX_CTAS");
    env_limit = env && atoi(env) > 0 ? atoi(env) : 0;
  }
  const int sms = device_sm_count();
  return env_limit > 0 ? std::min(sms, env_limit) : sms;
}

template <int BN, bool kCta2, int kEW, bool kAddSrc>
static int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mw, GemmKernelArgs& ka, cudaStream_t st) {
  using Cfg = GemmCfg<BN, kCta2, kEW>;
  static PerDeviceFlag attr_set;
  if (!attr_set.here()) {
    SEGMA_CUDA_OK(cudaFuncSetAttribute(gemm_tc5_kernel<BN, kCta2, kEW, kAddSrc>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set.here() = true;
  }
  ka.n_tiles = ceil_div(ka.n, BN);
  const int units_per_batch = kCta2 ? (ka.tiles_per_batch + 1) / 2 : ka.tiles_per_batch;
  const long long units = (long long)ka.batch * units_per_batch * ka.n_tiles;
  if (units > 0x7fffffffll) {
    set_last_error("segma_gemm_f16: too many tiles");
    return SEGMA_ERR_INVALID_ARGUMENT;
  }
  if (!kCta2) {
    const int grid = (int)std::min<long long>(units, gemm_cta_limit());
    gemm_tc5_kernel<BN, false, kEW, kAddSrc><<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(ma, mw, ka);
    return launch_status("gemm_tc5_kernel");
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * (unsigned)std::min<long long>(units, gemm_cta_limit() / 2));
  cfg.blockDim = dim3(Cfg::kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SEGMA_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc5_kernel<BN, true, kEW, kAddSrc>, ma, mw, ka));
  return SEGMA_OK;
}

}  // namespace segma

using namespace segma;

extern "C" {

int segma_gemm_f16(const segma_gemm_args* a, void* stream) {
  SEGMA_REQUIRE(a != nullptr, "segma_gemm_f16: NULL args");
  SEGMA_REQUIRE(a->batch >= 0 && a->rows_per_batch >= 0, "segma_gemm_f16: negative shape");
  if (a->batch == 0 || a->rows_per_batch == 0) return SEGMA_OK;
  SEGMA_REQUIRE(a->a && a->w && a->out, "segma_gemm_f16: NULL buffer");
  SEGMA_REQUIRE(a->n > 0 && a->n % 32 == 0, "segma_gemm_f16: n=%d must be a positive multiple of 32", a->n);
  SEGMA_REQUIRE(a->a_row_stride % 8 == 0 && a->a_batch_stride % 8 == 0,
                "segma_gemm_f16: A strides must be multiples of 8 elements (16 bytes)");
  SEGMA_REQUIRE((reinterpret_cast<uintptr_t>(a->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->w) & 15) == 0,
                "segma_gemm_f16: A and W must be 16-byte aligned");
  SEGMA_REQUIRE((reinterpret_cast<uintptr_t>(a->out) & 15) == 0 && a->ldo % 8 == 0,
                "segma_gemm_f16: out must be 16-byte aligned with ldo a multiple of 8");
  const int taps = a->conv_taps > 0 ? a->conv_taps : 1;
  const int s = a->conv_stride > 0 ? a->conv_stride : 1;
  SEGMA_REQUIRE(a->k > 0 && a->k % taps == 0, "segma_gemm_f16: k=%d must be a positive multiple of conv_taps=%d",
                a->k, taps);
  const int c = a->k / taps;  // channels per tap
  SEGMA_REQUIRE(c % 8 == 0, "segma_gemm_f16: channels per tap (%d) must be a multiple of 8", c);
  SEGMA_REQUIRE(taps == 1 || s == 1 || c % kBK == 0,
                "segma_gemm_f16: strided conv needs channels per tap (%d) to be a multiple of 64", c);
  SEGMA_REQUIRE(a->a_row_stride >= (int64_t)(a->a_cols > 0 ? a->a_cols : c), "segma_gemm_f16: a_row_stride smaller than the row");
  if (a->add_src) SEGMA_REQUIRE(a->add_batch_rows >= 0, "segma_gemm_f16: add_batch_rows must be non-negative");

  GemmKernelArgs ka{};
  ka.batch = a->batch;
  ka.rows_per_batch = a->rows_per_batch;
  ka.tiles_per_batch = ceil_div(a->rows_per_batch, kBM);
  ka.n = a->n;
  ka.taps = taps;
  ka.kb_per_tap = ceil_div(c, kBK);
  ka.conv_stride = s;
  ka.a_tap_elems = a->a_cols > 0 ? a->a_cols : c;
  ka.w_tap_elems = c;
  ka.a_ntile_off = a->a_col_per_ntile;
  ka.bias = a->bias;
  ka.add_src = a->add_src;
  ka.add_batch_rows = a->add_batch_rows;
  ka.out = a->out;
  ka.out_batch_rows = a->out_batch_rows;
  ka.out_row_offset = a->out_row_offset;
  ka.ldo = a->ldo;
  ka.flags = a->flags;
  {  // A/B switch for the epilogue route of fp16 outputs without a residual (read per call: tools flip it in-process)
    const char* env = getenv("SEGMA_GEMM_EPI_REGS");
    if (env ? atoi(env) != 0 : true) ka.flags |= kFlagEpiRegs;
  }

  // A map: stride-s convolutions view s consecutive input rows as one map row of s*c channels
  const int in_rows = a->a_rows_per_batch > 0 ? a->a_rows_per_batch : a->rows_per_batch;
  SEGMA_REQUIRE(in_rows % s == 0, "segma_gemm_f16: a_rows_per_batch=%d must be a multiple of conv_stride=%d", in_rows, s);
  CUtensorMap ma, mw;
  {
    const int row_cols = a->a_cols > 0 ? a->a_cols : c;
    uint64_t dims[3] = {(uint64_t)s * row_cols, (uint64_t)(in_rows / s), (uint64_t)a->batch};
    uint64_t strides[3] = {1, (uint64_t)a->a_row_stride * s, (uint64_t)a->a_batch_stride};
    if (a->batch == 1 && strides[2] == 0) strides[2] = strides[1] * dims[1];
    int rc = make_f16_map(&ma, a->a, 3, dims, strides, kBM);
    if (rc != SEGMA_OK) return rc;
  }
  int bn = 256;
  bool cta2 = false;
  if (a->n % 256 != 0 && a->n < 512) bn = 128;
  if (a->n % 256 != 0 && a->n % 128 != 0 && a->n % 192 == 0) bn = 192;
  if (a->force_bn) {
    SEGMA_REQUIRE(a->force_bn == 128 || a->force_bn == 192 || a->force_bn == 256, "segma_gemm_f16: bad force_bn");
    bn = a->force_bn;
  }
  {
    uint64_t dims[2] = {(uint64_t)a->k, (uint64_t)a->n};
    uint64_t strides[2] = {1, (uint64_t)a->k};
    // CTA pairs (cta_group::2) for the 256-wide tile when there are at least two M tiles per batch to pair up
    static const int pair_mode = getenv("SEGMA_GEMM_2CTA") ? atoi(getenv("SEGMA_GEMM_2CTA")) : 1;
    cta2 = pair_mode != 0 && bn == 256 && ka.tiles_per_batch >= 2;
    int rc = make_f16_map(&mw, a->w, 2, dims, strides, cta2 ? bn / 2 : bn);
    if (rc != SEGMA_OK) return rc;
  }
  cudaStream_t st = (cudaStream_t)stream;
  // epilogues without a residual operand need few registers: 16 epilogue warps (4 per scheduler) hide the latencies
  // of the ALU-bound erf-GELU (fc1: 800 -> 975 TFLOP/s) and of the fp16 store path (QKV: 1130 -> 1190)
  static const int wide_epi = getenv("SEGMA_GEMM_EPI16") ? atoi(getenv("SEGMA_GEMM_EPI16")) : 2;
  const bool heavy = wide_epi != 0 && ((a->flags & SEGMA_GEMM_GELU) || wide_epi == 2) && a->add_src == nullptr;
  switch (bn) {
    case 128: return launch_gemm<128, false, 8, true>(ma, mw, ka, st);
    case 192: return launch_gemm<192, false, 8, true>(ma, mw, ka, st);
    default:
      if (cta2) return heavy ? launch_gemm<256, true, 16, false>(ma, mw, ka, st) : launch_gemm<256, true, 8, true>(ma, mw, ka, st);
      return launch_gemm<256, false, 8, true>(ma, mw, ka, st);
  }
}

}  // extern "C"
