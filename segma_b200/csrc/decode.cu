// Stitching of per-window frame logits and threshold + run-length decoding into interval tables.
//
// Replaces torch.concat (src/segma/inference.py:209-211), apply_thresholds (214-234) and the host
// NumPy/Python run extraction of create_intervals (237-263).  All three are HBM-bound streaming
// passes: one coalesced read of the logits, warp-shuffle/ballot run-boundary detection, a block
// offset scan, and a compacted int32 table write.
#include <climits>
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace segma {

// ---------------------------------------------------------------------------------------------
// stitch: gather-mean over covering windows
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stitch_kernel(const float* __restrict__ win, int n_windows, int F, int sf,
                                                      int tail_frames, int C, float* __restrict__ out,
                                                      long long n_frames) {
  const long long total = n_frames * C;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long g = idx / C;
    const int c = static_cast<int>(idx - g * C);
    // windows i with i*sf <= g < i*sf + F_i ; the tail is window n_windows with tail_frames frames
    long long lo = (g - F + sf) / sf;  // ceil((g - F + 1) / sf) for g-F+1 >= 0
    if (g - F + 1 <= 0) lo = 0;
    long long hi = g / sf;
    float acc = 0.f;
    int cnt = 0;
    for (long long i = lo; i <= hi; ++i) {
      const long long k = g - i * sf;
      if (i < n_windows) {
        acc += __ldg(win + (i * F + k) * C + c);
        ++cnt;
      } else if (i == n_windows && k < tail_frames) {
        acc += __ldg(win + ((long long)n_windows * F + k) * C + c);
        ++cnt;
      }
    }
    out[idx] = cnt > 0 ? acc / static_cast<float>(cnt) : 0.f;
  }
}

// ---------------------------------------------------------------------------------------------
// threshold + run-length decode
// ---------------------------------------------------------------------------------------------
struct DecodeParams {
  float thr[SEGMA_MAX_LABELS];
  int C;
  int mode;
};

constexpr int kDecodeBlock = 1024;  // frames (= threads) per block

__device__ __forceinline__ bool is_active(float x, float thr, int mode) {
  if (mode == SEGMA_DECODE_SIGMOID) {
    // fp32 sigmoid exactly as written in the reference: 1 / (1 + exp(-x)), strict '>'
    const float s = 1.0f / (1.0f + expf(-x));
    return s > thr;
  }
  return x > thr;
}

__device__ __forceinline__ uint32_t frame_bits(const float* __restrict__ logits, long long frame,
                                               const DecodeParams& p) {
  uint32_t bits = 0;
  const float* row = logits + frame * p.C;
  if (p.C == 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(row));
    bits |= is_active(v.x, p.thr[0], p.mode) ? 1u : 0u;
    bits |= is_active(v.y, p.thr[1], p.mode) ? 2u : 0u;
    bits |= is_active(v.z, p.thr[2], p.mode) ? 4u : 0u;
    bits |= is_active(v.w, p.thr[3], p.mode) ? 8u : 0u;
  } else {
    for (int c = 0; c < p.C; ++c) bits |= is_active(__ldg(row + c), p.thr[c], p.mode) ? (1u << c) : 0u;
  }
  return bits;
}

__global__ void __launch_bounds__(256) mask_kernel(const float* __restrict__ logits, long long n_frames,
                                                    DecodeParams p, uint8_t* __restrict__ mask) {
  const long long total = n_frames * p.C;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = static_cast<int>(idx % p.C);
    mask[idx] = is_active(__ldg(logits + idx), p.thr[c], p.mode) ? 1 : 0;
  }
}

// Block -> (file, first frame) lookup: block_offsets[f] = first block of file f.
__device__ __forceinline__ int find_file(const int* __restrict__ block_offsets, int n_files, int block) {
  int lo = 0, hi = n_files - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(block_offsets + mid) <= block) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// One thread per block: which file does the block belong to (computed once per call; the decode / hysteresis
// kernels then start with a single load instead of a chain of dependent global loads)
__global__ void block_file_kernel(const int* __restrict__ block_offsets, int n_files, int total_blocks,
                                  int* __restrict__ block_file) {
  const int blk = blockIdx.x * blockDim.x + threadIdx.x;
  if (blk < total_blocks) block_file[blk] = find_file(block_offsets, n_files, blk);
}

// Pass 1: per-frame activity bits (stored for pass 3) and per-(file,label,block) counts of run
// starts and run ends.  counts index = C*block_offsets[file] + c*nblk_file + local_block.
__global__ void __launch_bounds__(kDecodeBlock) decode_count_kernel(
    const float* __restrict__ logits, const long long* __restrict__ file_offsets,
    const int* __restrict__ block_offsets, int n_files, DecodeParams p, uint32_t* __restrict__ bits_out,
    int* __restrict__ start_counts, int* __restrict__ end_counts) {
  __shared__ uint32_t s_bits[kDecodeBlock + 2];
  __shared__ int s_cnt[2][SEGMA_MAX_LABELS];
  const int file = find_file(block_offsets, n_files, blockIdx.x);
  const int local_block = blockIdx.x - block_offsets[file];
  const int nblk_file = block_offsets[file + 1] - block_offsets[file];
  const long long f_begin = file_offsets[file], f_end = file_offsets[file + 1];
  const long long frame = f_begin + (long long)local_block * kDecodeBlock + threadIdx.x;
  const bool valid = frame < f_end;

  if (threadIdx.x < 2 * SEGMA_MAX_LABELS) (&s_cnt[0][0])[threadIdx.x] = 0;
  uint32_t bits = valid ? frame_bits(logits, frame, p) : 0u;
  s_bits[threadIdx.x + 1] = bits;
  if (threadIdx.x == 0) {
    const long long prev = frame - 1;
    s_bits[0] = (prev >= f_begin) ? frame_bits(logits, prev, p) : 0u;
    const long long next = f_begin + (long long)(local_block + 1) * kDecodeBlock;
    s_bits[kDecodeBlock + 1] = (next < f_end) ? frame_bits(logits, next, p) : 0u;
  }
  if (valid) bits_out[frame] = bits;
  __syncthreads();
  const uint32_t prev = s_bits[threadIdx.x], next = s_bits[threadIdx.x + 2];
  const uint32_t starts = bits & ~prev, ends = bits & ~next;
  for (int c = 0; c < p.C; ++c) {
    const unsigned bs = __ballot_sync(0xffffffffu, (starts >> c) & 1u);
    const unsigned be = __ballot_sync(0xffffffffu, (ends >> c) & 1u);
    if (lane_id() == 0) {
      if (bs) atomicAdd(&s_cnt[0][c], __popc(bs));
      if (be) atomicAdd(&s_cnt[1][c], __popc(be));
    }
  }
  __syncthreads();
  if (threadIdx.x < p.C) {
    const long long idx = (long long)p.C * block_offsets[file] + (long long)threadIdx.x * nblk_file + local_block;
    start_counts[idx] = s_cnt[0][threadIdx.x];
    end_counts[idx] = s_cnt[1][threadIdx.x];
  }
}

// Pass 2: exclusive scan (single block) of both count arrays in place; total -> *count.
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 8;
__global__ void __launch_bounds__(kScanThreads) decode_scan_kernel(int* __restrict__ a, int* __restrict__ b,
                                                                    long long n, int* __restrict__ count) {
  __shared__ int s_warp[2][32];
  __shared__ int s_carry[2];
  if (threadIdx.x == 0) s_carry[0] = s_carry[1] = 0;
  __syncthreads();
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  for (long long base = 0; base < n; base += (long long)kScanThreads * kScanItems) {
    int va[kScanItems], vb[kScanItems];
    int sa = 0, sb = 0;
    const long long t0 = base + (long long)threadIdx.x * kScanItems;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
      const long long j = t0 + i;
      va[i] = j < n ? a[j] : 0;
      vb[i] = j < n ? b[j] : 0;
      sa += va[i];
      sb += vb[i];
    }
    int ia = sa, ib = sb;  // inclusive warp scan of per-thread sums
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
      if (lane >= o) { ia += ta; ib += tb; }
    }
    if (lane == 31) { s_warp[0][warp] = ia; s_warp[1][warp] = ib; }
    __syncthreads();
    if (warp == 0) {
      int wa = s_warp[0][lane], wb = s_warp[1][lane];
      int xa = wa, xb = wb;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int ta = __shfl_up_sync(0xffffffffu, xa, o), tb = __shfl_up_sync(0xffffffffu, xb, o);
        if (lane >= o) { xa += ta; xb += tb; }
      }
      s_warp[0][lane] = xa - wa;  // exclusive warp offsets
      s_warp[1][lane] = xb - wb;
    }
    __syncthreads();
    int ea = s_carry[0] + s_warp[0][warp] + ia - sa;
    int eb = s_carry[1] + s_warp[1][warp] + ib - sb;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
      const long long j = t0 + i;
      if (j < n) { a[j] = ea; b[j] = eb; }
      ea += va[i];
      eb += vb[i];
    }
    __syncthreads();
    if (threadIdx.x == kScanThreads - 1) { s_carry[0] = ea; s_carry[1] = eb; }
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = s_carry[0];
}

// Parallel version of pass 2 for long inputs: per-tile sums, a scan of the tile sums, per-tile scans with offsets.
constexpr int kScanTile = kScanThreads * kScanItems;  // 8192 elements per block
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(const int* __restrict__ a, const int* __restrict__ b,
                                                                       long long n, int* __restrict__ tile_sums) {
  __shared__ int s_a[32], s_b[32];
  const long long t0 = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
  int sa = 0, sb = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    const long long j = t0 + i;
    if (j < n) { sa += a[j]; sb += b[j]; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sa += __shfl_xor_sync(0xffffffffu, sa, o);
    sb += __shfl_xor_sync(0xffffffffu, sb, o);
  }
  if (lane_id() == 0) { s_a[threadIdx.x >> 5] = sa; s_b[threadIdx.x >> 5] = sb; }
  __syncthreads();
  if (threadIdx.x < 32) {
    int xa = s_a[threadIdx.x], xb = s_b[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      xa += __shfl_xor_sync(0xffffffffu, xa, o);
      xb += __shfl_xor_sync(0xffffffffu, xb, o);
    }
    if (threadIdx.x == 0) { tile_sums[2 * blockIdx.x] = xa; tile_sums[2 * blockIdx.x + 1] = xb; }
  }
}

// one block: exclusive scan of the (a, b) tile sums in place; total of `a` -> *count
__global__ void __launch_bounds__(kScanThreads) scan_tile_offsets_kernel(int* __restrict__ tile_sums, int n_tiles,
                                                                          int* __restrict__ count) {
  __shared__ int s_warp[2][32];
  __shared__ int s_carry[2];
  if (threadIdx.x == 0) s_carry[0] = s_carry[1] = 0;
  __syncthreads();
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  for (int base = 0; base < n_tiles; base += kScanThreads) {
    const int i = base + threadIdx.x;
    const int va = i < n_tiles ? tile_sums[2 * i] : 0, vb = i < n_tiles ? tile_sums[2 * i + 1] : 0;
    int ia = va, ib = vb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
      if (lane >= o) { ia += ta; ib += tb; }
    }
    if (lane == 31) { s_warp[0][warp] = ia; s_warp[1][warp] = ib; }
    __syncthreads();
    if (warp == 0) {
      const int wa = s_warp[0][lane], wb = s_warp[1][lane];
      int xa = wa, xb = wb;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int ta = __shfl_up_sync(0xffffffffu, xa, o), tb = __shfl_up_sync(0xffffffffu, xb, o);
        if (lane >= o) { xa += ta; xb += tb; }
      }
      s_warp[0][lane] = xa - wa;
      s_warp[1][lane] = xb - wb;
    }
    __syncthreads();
    const int ea = s_carry[0] + s_warp[0][warp] + ia - va, eb = s_carry[1] + s_warp[1][warp] + ib - vb;
    if (i < n_tiles) { tile_sums[2 * i] = ea; tile_sums[2 * i + 1] = eb; }
    __syncthreads();
    if (threadIdx.x == kScanThreads - 1) { s_carry[0] = ea + va; s_carry[1] = eb + vb; }
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = s_carry[0];
}

// exclusive scan of one 8192-element tile of both arrays in place, starting from the tile's offsets
__global__ void __launch_bounds__(kScanThreads) scan_tiles_kernel(int* __restrict__ a, int* __restrict__ b, long long n,
                                                                   const int* __restrict__ tile_offsets) {
  __shared__ int s_warp[2][32];
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const long long t0 = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
  int va[kScanItems], vb[kScanItems];
  int sa = 0, sb = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    const long long j = t0 + i;
    va[i] = j < n ? a[j] : 0;
    vb[i] = j < n ? b[j] : 0;
    sa += va[i];
    sb += vb[i];
  }
  int ia = sa, ib = sb;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
    if (lane >= o) { ia += ta; ib += tb; }
  }
  if (lane == 31) { s_warp[0][warp] = ia; s_warp[1][warp] = ib; }
  __syncthreads();
  if (warp == 0) {
    const int wa = s_warp[0][lane], wb = s_warp[1][lane];
    int xa = wa, xb = wb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int ta = __shfl_up_sync(0xffffffffu, xa, o), tb = __shfl_up_sync(0xffffffffu, xb, o);
      if (lane >= o) { xa += ta; xb += tb; }
    }
    s_warp[0][lane] = xa - wa;
    s_warp[1][lane] = xb - wb;
  }
  __syncthreads();
  int ea = tile_offsets[2 * blockIdx.x] + s_warp[0][warp] + ia - sa;
  int eb = tile_offsets[2 * blockIdx.x + 1] + s_warp[1][warp] + ib - sb;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    const long long j = t0 + i;
    if (j < n) { a[j] = ea; b[j] = eb; }
    ea += va[i];
    eb += vb[i];
  }
}

// Pass 3: recompute boundary flags from the stored bits, rank them inside the block and write the
// (file, label, start, end) rows.  The k-th start and the k-th end of a (file,label) pair form one run.
__global__ void __launch_bounds__(kDecodeBlock) decode_write_kernel(
    const uint32_t* __restrict__ bits_in, const long long* __restrict__ file_offsets,
    const int* __restrict__ block_offsets, int n_files, int C, const int* __restrict__ start_base,
    const int* __restrict__ end_base, int32_t* __restrict__ table, long long capacity) {
  __shared__ uint32_t s_bits[kDecodeBlock + 2];
  __shared__ int s_warp[2][kDecodeBlock / 32];
  const int file = find_file(block_offsets, n_files, blockIdx.x);
  const int local_block = blockIdx.x - block_offsets[file];
  const int nblk_file = block_offsets[file + 1] - block_offsets[file];
  const long long f_begin = file_offsets[file], f_end = file_offsets[file + 1];
  const long long frame = f_begin + (long long)local_block * kDecodeBlock + threadIdx.x;
  const bool valid = frame < f_end;
  const uint32_t bits = valid ? bits_in[frame] : 0u;
  s_bits[threadIdx.x + 1] = bits;
  if (threadIdx.x == 0) {
    s_bits[0] = (frame - 1 >= f_begin) ? bits_in[frame - 1] : 0u;
    const long long next = f_begin + (long long)(local_block + 1) * kDecodeBlock;
    s_bits[kDecodeBlock + 1] = (next < f_end) ? bits_in[next] : 0u;
  }
  __syncthreads();
  const uint32_t starts = bits & ~s_bits[threadIdx.x], ends = bits & ~s_bits[threadIdx.x + 2];
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int rel = static_cast<int>(frame - f_begin);
  for (int c = 0; c < C; ++c) {
    const bool is_s = (starts >> c) & 1u, is_e = (ends >> c) & 1u;
    const unsigned bs = __ballot_sync(0xffffffffu, is_s), be = __ballot_sync(0xffffffffu, is_e);
    if (lane == 0) { s_warp[0][warp] = __popc(bs); s_warp[1][warp] = __popc(be); }
    __syncthreads();
    if (warp == 0) {
      int a = s_warp[0][lane], b = s_warp[1][lane];
      int xa = a, xb = b;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int ta = __shfl_up_sync(0xffffffffu, xa, o), tb = __shfl_up_sync(0xffffffffu, xb, o);
        if (lane >= o) { xa += ta; xb += tb; }
      }
      s_warp[0][lane] = xa - a;
      s_warp[1][lane] = xb - b;
    }
    __syncthreads();
    const long long cidx = (long long)C * block_offsets[file] + (long long)c * nblk_file + local_block;
    const unsigned lt = (1u << lane) - 1u;
    if (is_s) {
      const long long row = (long long)start_base[cidx] + s_warp[0][warp] + __popc(bs & lt);
      if (row < capacity) {
        table[row * 4 + 0] = file;
        table[row * 4 + 1] = c;
        table[row * 4 + 2] = rel * SEGMA_FRAME_SAMPLES;
      }
    }
    if (is_e) {
      const long long row = (long long)end_base[cidx] + s_warp[1][warp] + __popc(be & lt);
      if (row < capacity) table[row * 4 + 3] = (rel + 1) * SEGMA_FRAME_SAMPLES;
    }
    __syncthreads();
  }
}

// ---- tiling of the hysteresis kernels (C <= 8): 4 frames per thread, one activity byte per frame --------------
constexpr int kFastThreads = 256;
constexpr int kFastFrames = 4;                             // consecutive frames per thread
constexpr int kFastBlock = kFastThreads * kFastFrames;     // 1024 frames per block (== kDecodeBlock)

// ---- bit-plane path for C <= 8 labels: one warp per 1024-frame block ----------------------------------------
// Count pass: lane l reads frame 32 k + l of the block (one coalesced 128-bit load per frame row for C = 4), a ballot
// per label turns 32 frames into one word, and lane k keeps the words of step k: after 32 steps lane k holds, for every
// label, the activity of frames [32 k, 32 k + 32) as a bit mask.  Run starts / ends are then word operations
// (w & ~(w << 1 | carry)), their counts population counts, and the block totals one packed warp reduction per label:
// about half an instruction per frame, so the pass runs at the speed of the logits read.  The planes (C words per
// lane, 128 bytes per label and block) are kept for the write pass, which never touches the logits again: it recomputes
// the boundary masks from the planes, ranks them with one packed warp scan per label and walks the set bits.
constexpr int kPlaneWarps = 8;      // blocks (of 1024 frames) per CTA
constexpr int kPlaneScanTile = 2048;  // block counts per scan CTA
constexpr int kPlaneMaxLabels = 8;

struct PlaneEdges {
  uint32_t starts, ends;
};
__device__ __forceinline__ PlaneEdges plane_edges(uint32_t w, uint32_t prev_bit, uint32_t next_bit, int lane) {
  const uint32_t up = __shfl_up_sync(0xffffffffu, w, 1), dn = __shfl_down_sync(0xffffffffu, w, 1);
  const uint32_t carry = lane == 0 ? prev_bit : up >> 31;
  const uint32_t nxt = lane == 31 ? next_bit : dn & 1u;
  PlaneEdges e;
  e.starts = w & ~((w << 1) | carry);
  e.ends = w & ~((w >> 1) | (nxt << 31));
  return e;
}

// totals of both count arrays per tile of 2048 counts
__global__ void __launch_bounds__(256) plane_tile_sums_kernel(const int* __restrict__ a, const int* __restrict__ b,
                                                              long long n, int* __restrict__ tile_sums) {
  __shared__ int s_warp[2][8];
  const long long t0 = (long long)blockIdx.x * kPlaneScanTile + threadIdx.x * 8;
  int sa = 0, sb = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (t0 + i < n) { sa += a[t0 + i]; sb += b[t0 + i]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); }
  if (lane_id() == 0) { s_warp[0][threadIdx.x >> 5] = sa; s_warp[1][threadIdx.x >> 5] = sb; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int ta = 0, tb = 0;
    for (int w = 0; w < 8; ++w) { ta += s_warp[0][w]; tb += s_warp[1][w]; }
    tile_sums[2 * blockIdx.x] = ta;
    tile_sums[2 * blockIdx.x + 1] = tb;
  }
}

// Exclusive scan of both count arrays in place, one CTA per tile of 2048 counts.  Every CTA sums the totals of the tiles before it (a few hundred values even for 1000 h of audio),
// so there is no separate offsets kernel and no inter-CTA dependency.  The last CTA writes the number of intervals.
__global__ void __launch_bounds__(256) plane_scan_kernel(int* __restrict__ a, int* __restrict__ b, long long n,
                                                         const int* __restrict__ tile_sums, int n_tiles,
                                                         int* __restrict__ count) {
  __shared__ int s_warp[2][8];
  __shared__ int s_base[2];
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  int pa = 0, pb = 0;
  for (int t = threadIdx.x; t < (int)blockIdx.x; t += 256) { pa += tile_sums[2 * t]; pb += tile_sums[2 * t + 1]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { pa += __shfl_xor_sync(0xffffffffu, pa, o); pb += __shfl_xor_sync(0xffffffffu, pb, o); }
  if (lane == 0) { s_warp[0][warp] = pa; s_warp[1][warp] = pb; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int ta = 0, tb = 0;
    for (int w = 0; w < 8; ++w) { ta += s_warp[0][w]; tb += s_warp[1][w]; }
    s_base[0] = ta;
    s_base[1] = tb;
    if ((int)blockIdx.x == n_tiles - 1) *count = ta + tile_sums[2 * blockIdx.x];
  }
  __syncthreads();
  // 8 consecutive counts per thread (two 128-bit accesses per array)
  const long long t0 = (long long)blockIdx.x * kPlaneScanTile + threadIdx.x * 8;
  int va[8], vb[8];
  int sa = 0, sb = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    va[i] = t0 + i < n ? a[t0 + i] : 0;
    vb[i] = t0 + i < n ? b[t0 + i] : 0;
    sa += va[i];
    sb += vb[i];
  }
  int ia = sa, ib = sb;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
    if (lane >= o) { ia += ta; ib += tb; }
  }
  __syncthreads();  // s_warp is reused
  if (lane == 31) { s_warp[0][warp] = ia; s_warp[1][warp] = ib; }
  __syncthreads();
  int ra = s_base[0] + ia - sa, rb = s_base[1] + ib - sb;
  for (int w = 0; w < warp; ++w) { ra += s_warp[0][w]; rb += s_warp[1][w]; }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (t0 + i < n) { a[t0 + i] = ra; b[t0 + i] = rb; }
    ra += va[i];
    rb += vb[i];
  }
}

// kC: compile-time label count (4 = segma's default label set, one 128-bit load per frame), 0 = read it from p.C
template <bool kWrite, int kC>
__global__ void __launch_bounds__(kPlaneWarps * 32) decode_plane_kernel(
    const float* __restrict__ logits, const long long* __restrict__ file_offsets,
    const int* __restrict__ block_offsets, int n_files, DecodeParams p, int total_blocks,
    const uint8_t* __restrict__ act, uint32_t* __restrict__ planes, int* __restrict__ start_counts,
    int* __restrict__ end_counts, int32_t* __restrict__ table, long long capacity) {
  const int lane = lane_id();
  const int blk = blockIdx.x * kPlaneWarps + (threadIdx.x >> 5);
  if (blk >= total_blocks) return;  // warp-uniform
  const int C = kC ? kC : p.C;
  constexpr int kMaxC = kC ? kC : kPlaneMaxLabels;
  // file of this block: last f with block_offsets[f] <= blk (empty files share their successor's offset)
  int file = 0;
  for (int lo = 0, hi = n_files; lo < hi;) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(block_offsets + mid + 1) <= blk) lo = mid + 1; else hi = mid;
    file = lo;
  }
  const int first_blk = block_offsets[file];
  const int local_block = blk - first_blk;
  const int nblk_file = block_offsets[file + 1] - first_blk;
  const long long f_begin = file_offsets[file], f_end = file_offsets[file + 1];
  const long long base = f_begin + (long long)local_block * kDecodeBlock;
  const long long cidx0 = (long long)C * first_blk + local_block;  // + c * nblk_file
  uint32_t* my_planes = planes + (size_t)blk * C * 32;

  uint32_t w[kMaxC];
  uint32_t prev_bits = 0, next_bits = 0;  // bit c: label c active in the frame before / after the block
  if (!kWrite) {
#pragma unroll
    for (int c = 0; c < kMaxC; ++c) w[c] = 0;
    const int n_here = (int)min((long long)kDecodeBlock, f_end - base);  // frames of this block that exist
    if (kC == 4 && p.mode == SEGMA_DECODE_LOGIT && act == nullptr) {  // the product's configuration: four compares and ballots per step
      const float4* rows = reinterpret_cast<const float4*>(logits) + base;
      const float t0 = p.thr[0], t1 = p.thr[1], t2 = p.thr[2], t3 = p.thr[3];
#pragma unroll 1
      for (int k0 = 0; k0 < 32; k0 += 8) {  // eight independent 128-bit loads in flight per lane, then their ballots
        if (32 * k0 >= n_here) break;  // warp-uniform
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[j] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
          if (32 * (k0 + j) + lane < n_here) v[j] = __ldg(rows + 32 * (k0 + j) + lane);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t b0 = __ballot_sync(0xffffffffu, v[j].x > t0), b1 = __ballot_sync(0xffffffffu, v[j].y > t1);
          const uint32_t b2 = __ballot_sync(0xffffffffu, v[j].z > t2), b3 = __ballot_sync(0xffffffffu, v[j].w > t3);
          if (lane == k0 + j) { w[0] = b0; w[1] = b1; w[2] = b2; w[3] = b3; }
        }
      }
    } else {
#pragma unroll 4
      for (int k = 0; k < 32; ++k) {
        if (32 * k >= n_here) break;  // warp-uniform
        // act: activity bytes resolved by the hysteresis kernels instead of thresholded logits
        uint32_t bits = 0;
        if (32 * k + lane < n_here) bits = act ? act[base + 32 * k + lane] : frame_bits(logits, base + 32 * k + lane, p);
#pragma unroll
        for (int c = 0; c < kMaxC; ++c) {
          if (c < C) {
            const uint32_t word = __ballot_sync(0xffffffffu, (bits >> c) & 1u);
            if (lane == k) w[c] = word;
          }
        }
      }
    }
    if (base - 1 >= f_begin) prev_bits = act ? act[base - 1] : frame_bits(logits, base - 1, p);  // same address in every lane
    if (base + kDecodeBlock < f_end) next_bits = act ? act[base + kDecodeBlock] : frame_bits(logits, base + kDecodeBlock, p);
  } else {
#pragma unroll
    for (int c = 0; c < kMaxC; ++c) {
      w[c] = 0;
      if (c < C) {
        w[c] = my_planes[c * 32 + lane];
        // neighbouring blocks of the same file: last bit of the previous one, first bit of the next one
        if (local_block > 0) prev_bits |= (__ldg(my_planes - C * 32 + c * 32 + 31) >> 31) << c;
        if (local_block + 1 < nblk_file) next_bits |= (__ldg(my_planes + C * 32 + c * 32) & 1u) << c;
      }
    }
  }

  const int rel0 = local_block * kDecodeBlock + 32 * lane;
  const unsigned cap = capacity > 0x7fffffffll ? 0x7fffffffu : static_cast<unsigned>(capacity);
#pragma unroll
  for (int c = 0; c < kMaxC; ++c) {
    if (c >= C) break;
    const PlaneEdges e = plane_edges(w[c], (prev_bits >> c) & 1u, (next_bits >> c) & 1u, lane);
    // starts in the low half, ends in the high half: a block holds at most 512 of each
    const uint32_t cnt = static_cast<uint32_t>(__popc(e.starts)) | (static_cast<uint32_t>(__popc(e.ends)) << 16);
    if (!kWrite) {
      my_planes[c * 32 + lane] = w[c];
      uint32_t tot = cnt;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
      if (lane == 0) {
        const long long idx = cidx0 + (long long)c * nblk_file;
        const int ns = static_cast<int>(tot & 0xffffu), ne = static_cast<int>(tot >> 16);
        start_counts[idx] = ns;
        end_counts[idx] = ne;
      }
    } else {
      uint32_t x = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += t;
      }
      x -= cnt;  // exclusive
      unsigned row_s = static_cast<unsigned>(__ldg(start_counts + cidx0 + (long long)c * nblk_file)) + (x & 0xffffu);
      unsigned row_e = static_cast<unsigned>(__ldg(end_counts + cidx0 + (long long)c * nblk_file)) + (x >> 16);
      for (uint32_t ms = e.starts, me = e.ends; ms | me; ms &= ms - 1, me &= me - 1, ++row_s, ++row_e) {
        if (ms && row_s < cap) {
          int* t = table + 4ll * row_s;
          t[0] = file;
          t[1] = c;
          t[2] = (rel0 + __ffs(ms) - 1) * SEGMA_FRAME_SAMPLES;
        }
        if (me && row_e < cap) table[4ll * row_e + 3] = (rel0 + __ffs(me)) * SEGMA_FRAME_SAMPLES;
      }
    }
  }
}

// ---- hysteresis (onset / offset thresholds): an extension that uses the `upper_bound` the reference carries in
// its threshold dict but never reads (src/segma/inference.py:308-312).  A label switches on when the logit
// exceeds the onset cut, off when it is at or below the offset cut, and otherwise keeps its state: a scan of
// (set, value) pairs, composed within threads, warps, blocks and -- through one tiny serial pass per file --
// across blocks.  With onset == offset it reduces to the plain threshold rule.
struct HystParams {
  float hi[8], lo[8];
  int C;
};

struct SetVal {
  uint32_t set, val;  // per-label bits: state is forced (set) to `val`
};
__device__ __forceinline__ SetVal compose(SetVal earlier, SetVal later) {
  SetVal r;
  r.set = earlier.set | later.set;
  r.val = (earlier.val & ~later.set) | (later.val & later.set);
  return r;
}

__device__ __forceinline__ uint32_t frame_codes(const float* __restrict__ logits, long long frame, const HystParams& p) {
  uint32_t hi = 0, lo = 0;
  const float* row = logits + frame * p.C;
  for (int c = 0; c < p.C; ++c) {
    const float x = __ldg(row + c);
    hi |= (x > p.hi[c]) ? (1u << c) : 0u;
    lo |= (x > p.lo[c]) ? (1u << c) : 0u;
  }
  return hi | (lo << 8);
}
__device__ __forceinline__ SetVal code_setval(uint32_t code, uint32_t label_mask) {
  const uint32_t hi = code & 0xffu, lo = (code >> 8) & 0xffu;
  SetVal r;
  r.set = (hi | ~lo) & label_mask;  // above onset -> on, at/below offset -> off
  r.val = hi & label_mask;
  return r;
}

// pass H1: per-frame (hi, lo) codes and the (set, val) summary of every 1024-frame block
__global__ void __launch_bounds__(kFastThreads) hyst_codes_kernel(
    const float* __restrict__ logits, const long long* __restrict__ file_offsets,
    const int* __restrict__ block_offsets, const int* __restrict__ block_file, HystParams p,
    uint16_t* __restrict__ codes, uint16_t* __restrict__ block_summary) {
  __shared__ SetVal s_w[kFastThreads / 32];
  const int file = __ldg(block_file + blockIdx.x);
  const int local_block = blockIdx.x - block_offsets[file];
  const long long f_begin = file_offsets[file], f_end = file_offsets[file + 1];
  const long long frame0 = f_begin + (long long)local_block * kFastBlock + threadIdx.x * kFastFrames;
  const uint32_t label_mask = (1u << p.C) - 1u;
  SetVal sv{0u, 0u};
#pragma unroll
  for (int i = 0; i < kFastFrames; ++i) {
    if (frame0 + i < f_end) {
      const uint32_t code = frame_codes(logits, frame0 + i, p);
      codes[frame0 + i] = static_cast<uint16_t>(code);
      sv = compose(sv, code_setval(code, label_mask));
    }
  }
  const int lane = lane_id(), warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {  // inclusive scan in lane order; lane 31 ends with the warp total
    SetVal up;
    up.set = __shfl_up_sync(0xffffffffu, sv.set, o);
    up.val = __shfl_up_sync(0xffffffffu, sv.val, o);
    if (lane >= o) sv = compose(up, sv);
  }
  if (lane == 31) s_w[warp] = sv;
  __syncthreads();
  if (threadIdx.x == 0) {
    SetVal t = s_w[0];
    for (int w = 1; w < kFastThreads / 32; ++w) t = compose(t, s_w[w]);
    block_summary[blockIdx.x] = static_cast<uint16_t>((t.set & 0xffu) | ((t.val & 0xffu) << 8));
  }
}

// pass H2: state entering every block (one thread per file walks its blocks)
__global__ void hyst_carry_kernel(const uint16_t* __restrict__ block_summary, const int* __restrict__ block_offsets,
                                  int n_files, uint8_t* __restrict__ carry_in) {
  const int file = blockIdx.x * blockDim.x + threadIdx.x;
  if (file >= n_files) return;
  uint32_t state = 0;
  for (int b = block_offsets[file]; b < block_offsets[file + 1]; ++b) {
    carry_in[b] = static_cast<uint8_t>(state);
    const uint32_t s = block_summary[b] & 0xffu, v = block_summary[b] >> 8;
    state = (state & ~s) | (v & s);
  }
}

// pass H3: resolve the state of every frame -> the activity byte the decode passes consume
__global__ void __launch_bounds__(kFastThreads) hyst_resolve_kernel(
    const uint16_t* __restrict__ codes, const long long* __restrict__ file_offsets,
    const int* __restrict__ block_offsets, const int* __restrict__ block_file, int C,
    const uint8_t* __restrict__ carry_in, uint8_t* __restrict__ bits) {
  __shared__ SetVal s_w[kFastThreads / 32];
  const int file = __ldg(block_file + blockIdx.x);
  const int local_block = blockIdx.x - block_offsets[file];
  const long long f_begin = file_offsets[file], f_end = file_offsets[file + 1];
  const long long frame0 = f_begin + (long long)local_block * kFastBlock + threadIdx.x * kFastFrames;
  const uint32_t label_mask = (1u << C) - 1u;
  SetVal own[kFastFrames];
  SetVal sv{0u, 0u};
#pragma unroll
  for (int i = 0; i < kFastFrames; ++i) {
    own[i] = SetVal{0u, 0u};
    if (frame0 + i < f_end) own[i] = code_setval(codes[frame0 + i], label_mask);
    sv = compose(sv, own[i]);
  }
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  SetVal inc = sv;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    SetVal up;
    up.set = __shfl_up_sync(0xffffffffu, inc.set, o);
    up.val = __shfl_up_sync(0xffffffffu, inc.val, o);
    if (lane >= o) inc = compose(up, inc);
  }
  if (lane == 31) s_w[warp] = inc;
  SetVal excl;  // composition of the lanes before this one
  excl.set = __shfl_up_sync(0xffffffffu, inc.set, 1);
  excl.val = __shfl_up_sync(0xffffffffu, inc.val, 1);
  if (lane == 0) excl = SetVal{0u, 0u};
  __syncthreads();
  SetVal before{0u, 0u};
  for (int w = 0; w < warp; ++w) before = compose(before, s_w[w]);
  before = compose(before, excl);
  uint32_t state = carry_in[blockIdx.x];
  state = (state & ~before.set) | before.val;
#pragma unroll
  for (int i = 0; i < kFastFrames; ++i) {
    state = (state & ~own[i].set) | own[i].val;
    if (frame0 + i < f_end) bits[frame0 + i] = static_cast<uint8_t>(state);
  }
}

// ---- interval-table post-processing: merge runs of the same (file, label) separated by at most max_gap samples
// (max_gap = 0 merges adjacent / overlapping intervals like the reference's Intervals struct,
// src/segma/structs/interval.py:19-34), then drop intervals shorter than min_dur samples.  Tables are small
// (KBs to a few MB): one block, chunked scans.
constexpr int kPostThreads = 1024;

__device__ __forceinline__ int block_incl_scan(int v, int* s_warp, int lane, int warp) {
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += t;
  }
  if (lane == 31) s_warp[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int w = s_warp[lane], y = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, y, o);
      if (lane >= o) y += t;
    }
    s_warp[lane] = y - w;
  }
  __syncthreads();
  const int r = x + s_warp[warp];
  __syncthreads();
  return r;
}

// Segmented running maximum of the interval ends: (flag, val) pairs under (a then b) -> (a.flag | b.flag,
// b.flag ? b.val : max(a.val, b.val)); flag marks the first row of a (file, label) segment.
struct SegMax { int flag, val; };
__device__ __forceinline__ SegMax segmax(SegMax a, SegMax b) {
  return SegMax{a.flag | b.flag, b.flag ? b.val : max(a.val, b.val)};
}

__device__ __forceinline__ SegMax block_incl_segmax(SegMax v, int* s_flag, int* s_val, int lane, int warp) {
  SegMax x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    SegMax t{__shfl_up_sync(0xffffffffu, x.flag, o), __shfl_up_sync(0xffffffffu, x.val, o)};
    if (lane >= o) x = segmax(t, x);
  }
  if (lane == 31) { s_flag[warp] = x.flag; s_val[warp] = x.val; }
  __syncthreads();
  if (warp == 0) {
    SegMax y{s_flag[lane], s_val[lane]};
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      SegMax t{__shfl_up_sync(0xffffffffu, y.flag, o), __shfl_up_sync(0xffffffffu, y.val, o)};
      if (lane >= o) y = segmax(t, y);
    }
    s_flag[lane] = y.flag;  // inclusive over warps 0..lane
    s_val[lane] = y.val;
  }
  __syncthreads();
  if (warp > 0) x = segmax(SegMax{s_flag[warp - 1], s_val[warp - 1]}, x);
  __syncthreads();
  return x;
}

// Rows sorted by (file, label, start).  A row opens a new interval when its (file, label) differs from the previous
// row's or its start lies more than max_gap past the largest end seen so far in that (file, label) -- the reference's
// `s <= ret[-1][1]` test (interval.py:26-31) generalised by the gap -- so overlapping and nested rows merge too; the
// merged end is that running maximum.
__global__ void __launch_bounds__(kPostThreads) merge_intervals_kernel(const int4* __restrict__ in, long long n,
                                                                       int max_gap, int4* __restrict__ merged,
                                                                       int* __restrict__ n_merged) {
  __shared__ int s_warp[32], s_flag[32], s_val[32];
  __shared__ int s_incl[kPostThreads];
  __shared__ int s_carry, s_carry_end;
  if (threadIdx.x == 0) { s_carry = 0; s_carry_end = INT_MIN; }
  __syncthreads();
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  for (long long base = 0; base < n; base += kPostThreads) {
    const long long i = base + threadIdx.x;
    int4 row = make_int4(0, 0, 0, INT_MIN);
    int key_head = 0;
    if (i < n) {
      row = in[i];
      if (i == 0) {
        key_head = 1;
      } else {
        const int4 pr = in[i - 1];
        key_head = (pr.x != row.x || pr.y != row.y) ? 1 : 0;
      }
    }
    SegMax m = block_incl_segmax(SegMax{key_head, row.w}, s_flag, s_val, lane, warp);
    const int carry_end = s_carry_end;
    if (!m.flag) m.val = max(m.val, carry_end);  // same (file, label) as the last row of the previous chunk
    s_incl[threadIdx.x] = m.val;
    __syncthreads();
    int head = 0, last = 0;
    if (i < n) {
      const int before = threadIdx.x > 0 ? s_incl[threadIdx.x - 1] : carry_end;
      head = (key_head || (long long)row.z - before > max_gap) ? 1 : 0;
      if (i == n - 1) {
        last = 1;
      } else {
        const int4 nx = in[i + 1];
        last = (nx.x != row.x || nx.y != row.y || (long long)nx.z - m.val > max_gap) ? 1 : 0;
      }
    }
    const int g = s_carry + block_incl_scan(head, s_warp, lane, warp) - 1;  // group index of row i
    if (i < n) {
      if (head) { merged[g].x = row.x; merged[g].y = row.y; merged[g].z = row.z; }
      if (last) merged[g].w = m.val;
    }
    __syncthreads();
    if (threadIdx.x == kPostThreads - 1) { s_carry = g + 1; s_carry_end = m.val; }
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_merged = s_carry;
}

__global__ void __launch_bounds__(kPostThreads) filter_intervals_kernel(const int4* __restrict__ merged,
                                                                        const int* __restrict__ n_merged, int min_dur,
                                                                        int4* __restrict__ out, long long capacity,
                                                                        int* __restrict__ count) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const long long n = *n_merged;
  for (long long base = 0; base < n; base += kPostThreads) {
    const long long i = base + threadIdx.x;
    int keep = 0;
    int4 row = make_int4(0, 0, 0, 0);
    if (i < n) {
      row = merged[i];
      keep = (row.w - row.z >= min_dur) ? 1 : 0;
    }
    const int pos = s_carry + block_incl_scan(keep, s_warp, lane, warp) - 1;
    SEGMA_DEV_ASSERT(!keep || (row.w >= row.z && pos >= 0));
    if (keep && pos < capacity) out[pos] = row;
    __syncthreads();
    if (threadIdx.x == kPostThreads - 1) s_carry = pos + 1;
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = s_carry;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct DecodeLayout {
  size_t bits_off, act_off, starts_off, ends_off, file_off, block_off, codes_off, summary_off, carry_off, bfile_off, tiles_off, total;
  long long n_count;
};

static DecodeLayout decode_layout(long long n_frames, int n_files, int C) {
  DecodeLayout L;
  const long long max_blocks = ceil_div_ll(n_frames, kDecodeBlock) + n_files;
  L.n_count = max_blocks * C;
  size_t off = 0;
  // one word per frame (generic path), one byte per frame (byte path) or C x 128 bytes per block (bit planes)
  L.bits_off = off; off = align_up(off + std::max(sizeof(uint32_t) * (size_t)n_frames, (size_t)max_blocks * C * 128), 256);
  L.act_off = off; off = align_up(off + (size_t)n_frames, 256);  // hysteresis: resolved activity bytes
  L.starts_off = off; off = align_up(off + sizeof(int) * (size_t)L.n_count, 256);
  L.ends_off = off; off = align_up(off + sizeof(int) * (size_t)L.n_count, 256);
  L.file_off = off; off = align_up(off + sizeof(long long) * (size_t)(n_files + 1), 256);
  L.block_off = off; off = align_up(off + sizeof(int) * (size_t)(n_files + 1), 256);
  L.codes_off = off; off = align_up(off + sizeof(uint16_t) * (size_t)n_frames, 256);
  L.summary_off = off; off = align_up(off + sizeof(uint16_t) * (size_t)max_blocks, 256);
  L.carry_off = off; off = align_up(off + (size_t)max_blocks, 256);
  L.bfile_off = off; off = align_up(off + sizeof(int) * (size_t)max_blocks, 256);
  L.tiles_off = off; off = align_up(off + 2 * sizeof(int) * (size_t)(L.n_count / 2048 + 2), 256);
  L.total = off;
  return L;
}

static int fill_params(DecodeParams& p, int C, const float* thr, int mode) {
  SEGMA_REQUIRE(C >= 1 && C <= SEGMA_MAX_LABELS, "n_labels must be in [1, %d], got %d", SEGMA_MAX_LABELS, C);
  SEGMA_REQUIRE(mode == SEGMA_DECODE_SIGMOID || mode == SEGMA_DECODE_LOGIT, "unknown decode mode %d", mode);
  SEGMA_REQUIRE(thr != nullptr, "thresholds is NULL");
  p.C = C;
  p.mode = mode;
  for (int c = 0; c < SEGMA_MAX_LABELS; ++c) p.thr[c] = c < C ? thr[c] : 0.f;
  return SEGMA_OK;
}

}  // namespace segma

using namespace segma;

extern "C" {

int segma_stitch(const float* window_logits, int n_windows, int frames_per_window, int step_frames, int tail_frames,
                 int n_labels, float* out, int64_t n_frames, void* stream) {
  SEGMA_REQUIRE(n_windows >= 0 && frames_per_window > 0 && step_frames > 0 && tail_frames >= 0 && n_labels > 0,
                "segma_stitch: bad geometry");
  SEGMA_REQUIRE(step_frames <= frames_per_window, "segma_stitch: step_frames %d leaves gaps (frames_per_window %d)",
                step_frames, frames_per_window);
  if (n_frames == 0) return SEGMA_OK;
  SEGMA_REQUIRE(window_logits && out, "segma_stitch: NULL buffer");
  const long long total = (long long)n_frames * n_labels;
  const int grid = (int)std::min<long long>(ceil_div_ll(total, 256), (long long)device_sm_count() * 16);
  stitch_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(window_logits, n_windows, frames_per_window, step_frames,
                                                        tail_frames, n_labels, out, n_frames);
  return launch_status("stitch_kernel");
}

int segma_threshold_mask(const float* logits, int64_t n_frames, int n_labels, const float* thresholds, int mode,
                         uint8_t* mask, void* stream) {
  DecodeParams p;
  int rc = fill_params(p, n_labels, thresholds, mode);
  if (rc != SEGMA_OK) return rc;
  if (n_frames == 0) return SEGMA_OK;
  SEGMA_REQUIRE(logits && mask, "segma_threshold_mask: NULL buffer");
  const long long total = (long long)n_frames * n_labels;
  const int grid = (int)std::min<long long>(ceil_div_ll(total, 256), (long long)device_sm_count() * 16);
  mask_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(logits, n_frames, p, mask);
  return launch_status("mask_kernel");
}

size_t segma_decode_workspace_bytes(int64_t n_frames, int n_files, int n_labels) {
  if (n_frames < 0 || n_files < 1 || n_labels < 1) return 0;
  return decode_layout(n_frames, n_files, n_labels).total;
}

static int decode_impl(const float* logits, const int64_t* file_offsets, int n_files, int n_labels,
                       const float* thresholds, const float* onset, int mode, int32_t* table, int64_t capacity,
                       int32_t* count, void* workspace, size_t workspace_bytes, void* stream) {
  DecodeParams p;
  int rc = fill_params(p, n_labels, thresholds, mode);
  if (rc != SEGMA_OK) return rc;
  SEGMA_REQUIRE(n_files >= 1 && file_offsets, "segma_decode_intervals: need at least one file");
  SEGMA_REQUIRE(count && workspace, "segma_decode_intervals: NULL count/workspace");
  SEGMA_REQUIRE(capacity >= 0 && (capacity == 0 || table), "segma_decode_intervals: NULL table");
  SEGMA_REQUIRE(file_offsets[0] == 0, "segma_decode_intervals: file_offsets[0] must be 0");
  cudaStream_t st = (cudaStream_t)stream;
  const long long n_frames = file_offsets[n_files];
  std::vector<int> block_offsets(n_files + 1, 0);
  for (int f = 0; f < n_files; ++f) {
    const long long len = file_offsets[f + 1] - file_offsets[f];
    SEGMA_REQUIRE(len >= 0, "segma_decode_intervals: file_offsets must be non-decreasing");
    SEGMA_REQUIRE(len * SEGMA_FRAME_SAMPLES < (1ll << 31), "segma_decode_intervals: file %d exceeds int32 samples", f);
    block_offsets[f + 1] = block_offsets[f] + (int)ceil_div_ll(len, kDecodeBlock);
  }
  const int total_blocks = block_offsets[n_files];
  const DecodeLayout L = decode_layout(n_frames, n_files, n_labels);
  SEGMA_REQUIRE(workspace_bytes >= L.total, "segma_decode_intervals: workspace too small (%zu < %zu)",
                workspace_bytes, L.total);
  if (total_blocks == 0) return check_cuda(cudaMemsetAsync(count, 0, sizeof(int32_t), st), "memset count");
  SEGMA_REQUIRE(logits, "segma_decode_intervals: NULL logits");
  char* ws = static_cast<char*>(workspace);
  uint32_t* bits = reinterpret_cast<uint32_t*>(ws + L.bits_off);
  int* starts = reinterpret_cast<int*>(ws + L.starts_off);
  int* ends = reinterpret_cast<int*>(ws + L.ends_off);
  long long* d_file = reinterpret_cast<long long*>(ws + L.file_off);
  int* d_block = reinterpret_cast<int*>(ws + L.block_off);
  static_assert(sizeof(long long) == sizeof(int64_t), "int64 layout");
  SEGMA_CUDA_OK(cudaMemcpyAsync(d_file, file_offsets, sizeof(int64_t) * (n_files + 1), cudaMemcpyHostToDevice, st));
  SEGMA_CUDA_OK(cudaMemcpyAsync(d_block, block_offsets.data(), sizeof(int) * (n_files + 1), cudaMemcpyHostToDevice, st));
  int* d_bfile = reinterpret_cast<int*>(ws + L.bfile_off);
  int* d_tiles = reinterpret_cast<int*>(ws + L.tiles_off);
  // pageable-source async copies are staged before returning, so block_offsets may go out of scope
  const bool fast = n_labels <= 8;  // bit planes (built from thresholded logits, or from the resolved hysteresis bytes)
  const uint8_t* act = nullptr;
  const long long n_counts = (long long)total_blocks * n_labels;
  // long inputs: two-kernel scan over tiles of 2048 block counts
  const int plane_tiles = (fast && n_counts > 4 * kScanTile) ? (int)ceil_div_ll(n_counts, kPlaneScanTile) : 0;
  if (!fast || onset) block_file_kernel<<<ceil_div(total_blocks, 256), 256, 0, st>>>(d_block, n_files, total_blocks, d_bfile);
  static_assert(kFastBlock == kDecodeBlock, "both paths tile files in blocks of 1024 frames");
  if (onset) {
    SEGMA_REQUIRE(fast && mode == SEGMA_DECODE_LOGIT, "hysteresis needs logit-domain cuts and at most 8 labels");
    HystParams hp;
    hp.C = n_labels;
    for (int c = 0; c < 8; ++c) {
      hp.hi[c] = c < n_labels ? onset[c] : 0.f;
      hp.lo[c] = c < n_labels ? thresholds[c] : 0.f;
      if (c < n_labels) SEGMA_REQUIRE(!(hp.hi[c] < hp.lo[c]), "hysteresis: onset cut below offset cut for label %d", c);
    }
    uint16_t* codes = reinterpret_cast<uint16_t*>(ws + L.codes_off);
    uint16_t* summary = reinterpret_cast<uint16_t*>(ws + L.summary_off);
    uint8_t* carry = reinterpret_cast<uint8_t*>(ws + L.carry_off);
    hyst_codes_kernel<<<total_blocks, kFastThreads, 0, st>>>(logits, d_file, d_block, d_bfile, hp, codes, summary);
    rc = launch_status("hyst_codes_kernel");
    if (rc != SEGMA_OK) return rc;
    hyst_carry_kernel<<<ceil_div(n_files, 128), 128, 0, st>>>(summary, d_block, n_files, carry);
    rc = launch_status("hyst_carry_kernel");
    if (rc != SEGMA_OK) return rc;
    uint8_t* act_out = reinterpret_cast<uint8_t*>(ws + L.act_off);
    hyst_resolve_kernel<<<total_blocks, kFastThreads, 0, st>>>(codes, d_file, d_block, d_bfile, n_labels, carry, act_out);
    act = act_out;
    rc = launch_status("hyst_resolve_kernel");
    if (rc != SEGMA_OK) return rc;
  }
  if (fast) {
    if (n_labels == 4)
      decode_plane_kernel<false, 4><<<ceil_div(total_blocks, kPlaneWarps), kPlaneWarps * 32, 0, st>>>(
          logits, d_file, d_block, n_files, p, total_blocks, act, bits, starts, ends, table, capacity);
    else
      decode_plane_kernel<false, 0><<<ceil_div(total_blocks, kPlaneWarps), kPlaneWarps * 32, 0, st>>>(
          logits, d_file, d_block, n_files, p, total_blocks, act, bits, starts, ends, table, capacity);
  } else {
    decode_count_kernel<<<total_blocks, kDecodeBlock, 0, st>>>(logits, d_file, d_block, n_files, p, bits, starts, ends);
  }
  rc = launch_status("decode count pass");
  if (rc != SEGMA_OK) return rc;
  if (plane_tiles) {
    plane_tile_sums_kernel<<<plane_tiles, 256, 0, st>>>(starts, ends, n_counts, d_tiles);
    plane_scan_kernel<<<plane_tiles, 256, 0, st>>>(starts, ends, n_counts, d_tiles, plane_tiles, count);
  } else if (n_counts <= 4 * kScanTile) {
    decode_scan_kernel<<<1, kScanThreads, 0, st>>>(starts, ends, n_counts, count);
  } else {  // long inputs: scan in parallel
    const int n_tiles = (int)ceil_div_ll(n_counts, kScanTile);
    scan_tile_sums_kernel<<<n_tiles, kScanThreads, 0, st>>>(starts, ends, n_counts, d_tiles);
    scan_tile_offsets_kernel<<<1, kScanThreads, 0, st>>>(d_tiles, n_tiles, count);
    scan_tiles_kernel<<<n_tiles, kScanThreads, 0, st>>>(starts, ends, n_counts, d_tiles);
  }
  rc = launch_status("decode scan pass");
  if (rc != SEGMA_OK) return rc;
  if (fast) {
    if (n_labels == 4)
      decode_plane_kernel<true, 4><<<ceil_div(total_blocks, kPlaneWarps), kPlaneWarps * 32, 0, st>>>(
          logits, d_file, d_block, n_files, p, total_blocks, nullptr, bits, starts, ends, table, capacity);
    else
      decode_plane_kernel<true, 0><<<ceil_div(total_blocks, kPlaneWarps), kPlaneWarps * 32, 0, st>>>(
          logits, d_file, d_block, n_files, p, total_blocks, nullptr, bits, starts, ends, table, capacity);
  } else {
    decode_write_kernel<<<total_blocks, kDecodeBlock, 0, st>>>(bits, d_file, d_block, n_files, n_labels, starts, ends,
                                                               table, capacity);
  }
  return launch_status("decode write pass");
}

int segma_decode_intervals(const float* logits, const int64_t* file_offsets, int n_files, int n_labels,
                           const float* thresholds, int mode, int32_t* table, int64_t capacity, int32_t* count,
                           void* workspace, size_t workspace_bytes, void* stream) {
  return decode_impl(logits, file_offsets, n_files, n_labels, thresholds, nullptr, mode, table, capacity, count,
                     workspace, workspace_bytes, stream);
}

int segma_decode_intervals_hysteresis(const float* logits, const int64_t* file_offsets, int n_files, int n_labels,
                                      const float* offset_cuts, const float* onset_cuts, int32_t* table,
                                      int64_t capacity, int32_t* count, void* workspace, size_t workspace_bytes,
                                      void* stream) {
  SEGMA_REQUIRE(onset_cuts != nullptr, "segma_decode_intervals_hysteresis: NULL onset cuts");
  return decode_impl(logits, file_offsets, n_files, n_labels, offset_cuts, onset_cuts, SEGMA_DECODE_LOGIT, table,
                     capacity, count, workspace, workspace_bytes, stream);
}

int segma_postprocess_intervals(const int32_t* table, int64_t n, int max_gap_samples, int min_duration_samples,
                                int32_t* scratch, int32_t* out, int64_t capacity, int32_t* counts, void* stream) {
  SEGMA_REQUIRE(n >= 0 && max_gap_samples >= 0 && min_duration_samples >= 0, "segma_postprocess_intervals: bad arguments");
  SEGMA_REQUIRE(counts, "segma_postprocess_intervals: NULL counts");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) return check_cuda(cudaMemsetAsync(counts, 0, 2 * sizeof(int32_t), st), "memset counts");
  SEGMA_REQUIRE(table && scratch && out, "segma_postprocess_intervals: NULL buffer");
  merge_intervals_kernel<<<1, kPostThreads, 0, st>>>(reinterpret_cast<const int4*>(table), n, max_gap_samples,
                                                    reinterpret_cast<int4*>(scratch), counts);
  int rc = launch_status("merge_intervals_kernel");
  if (rc != SEGMA_OK) return rc;
  filter_intervals_kernel<<<1, kPostThreads, 0, st>>>(reinterpret_cast<const int4*>(scratch), counts,
                                                     min_duration_samples, reinterpret_cast<int4*>(out), capacity,
                                                     counts + 1);
  return launch_status("filter_intervals_kernel");
}

}  // extern "C"
