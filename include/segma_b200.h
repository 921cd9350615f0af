/* libsegma_b200 -- C ABI of the B200-native sliding-window inference path of arxaqapi/segma.
 *
 * segma itself is pure Python and has no FFI: the drop-in boundary is its
 * Python prediction API (segma_b200/inference.py keeps the signatures of
 * /root/reference/src/segma/inference.py).  This header is the flat, torch-free layer underneath:
 * plain pointers and sizes, one `cudaStream_t` (passed as `void*`) per call, every buffer owned by
 * the caller, no hidden allocation or synchronisation unless stated.  Each entry point names the
 * reference code it replaces (paths relative to /root/reference, `site-packages/` = third-party).
 *
 * All functions return SEGMA_OK (0) or a negative error code; `segma_last_error()` holds the text.
 * Device pointers are marked [dev], host pointers [host].
 */
#ifndef SEGMA_B200_H
#define SEGMA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEGMA_API __attribute__((visibility("default")))

#define SEGMA_OK 0
#define SEGMA_ERR_INVALID_ARGUMENT (-1)
#define SEGMA_ERR_CUDA (-2)
#define SEGMA_ERR_UNSUPPORTED (-3)

#define SEGMA_MAX_LABELS 32
#define SEGMA_FRAME_SAMPLES 320
#define SEGMA_MEL_BINS 80
#define SEGMA_MEL_FRAMES 3000

/* ---- library ------------------------------------------------------------------------------ */
SEGMA_API const char* segma_last_error(void);
SEGMA_API int segma_version(void);
/* SEGMA_OK iff the current CUDA device is compute capability 10.x (the kernels are sm_100a only). */
SEGMA_API int segma_device_check(void);
SEGMA_API int segma_sm_count(void);

/* ---- audio staging ---------------------------------------------------------------------------
 * Widening of native-width PCM to the float32 samples the reference's decoder hands to the model
 * (src/segma/utils/io.py:30-47, torchcodec: int16 / 2^15, int32 / 2^31), on the device, so a file crosses
 * PCIe in its stored width.  src [dev], dst [dev] (n) fp32.
 */
#define SEGMA_PCM_S16 0
#define SEGMA_PCM_S32 1
#define SEGMA_PCM_F32 2
SEGMA_API int segma_pcm_to_f32(const void* src, int format, int64_t n, float* dst, void* stream);

/* ---- windowing + Whisper log-mel front end --------------------------------------------------
 * Replaces `sub_audio_t.unfold(0, 64000, 63680)` (src/segma/inference.py:148-152) fused with
 * `audio_preparation_hook` -> WhisperFeatureExtractor._torch_extract_fbank_features
 * (src/segma/models/whisper/hydra.py:197-201;
 * site-packages/transformers/models/whisper/feature_extraction_whisper.py:135-164):
 * window i = pcm[i*step, i*step+win_len) zero-padded to 30 s, STFT(400, hop 160, periodic Hann,
 * centre/reflect), power, 80 slaney mel bins, log10, per-window max-8 clamp, (x+4)/4.
 * Samples at or beyond pcm_len read as zero (a tail window is n_windows=1 with its own win_len).
 *   out_f32  [dev] (n_windows, 80, 3000) fp32, or NULL
 *   out_tm   [dev] (n_windows, 3002, 80) fp16 time-major with one zero row before and after each
 *            window (the layout the conv-stem implicit GEMM reads), or NULL
 *   scratch  [dev] segma_logmel_scratch_bytes(n_windows, win_len) bytes, 128-byte aligned (cudaMalloc is)
 */
SEGMA_API size_t segma_logmel_scratch_bytes(int n_windows, int win_len);
SEGMA_API int segma_logmel(const float* pcm, int64_t pcm_len, int n_windows, int win_len, int64_t step, float* out_f32,
                 void* out_tm, void* scratch, void* stream);
/* Replace the built-in slaney filterbank by a dense fp32 (201, 80) matrix [host]. Synchronous. */
SEGMA_API int segma_logmel_set_filters(const float* mel_201x80);
/* Copy the active dense (201, 80) filterbank to [host] memory. */
SEGMA_API int segma_logmel_get_filters(float* mel_201x80);

/* ---- wav2vec2 / HuBERT / WavLM waveform front end, layer 0 ------------------------------------
 * Conv1d(1, C, k=10, s=5, no bias) -> GroupNorm(C groups: per (window, channel) over time, eps 1e-5, biased
 * variance) -> exact GELU (site-packages/torchaudio/models/wav2vec2/components.py:77-99), fused with the
 * windowing of src/segma/inference.py:148-152.  out: fp16 time-major (n_windows, out_rows, C), rows at or
 * beyond the (win_len-10)/5+1 conv outputs are written as zeros; scale_shift: (n_windows, C, 2) fp32 scratch.
 * C must be a multiple of 4 (a thread owns four channels).
 */
SEGMA_API int segma_w2v2_layer0(const float* pcm, int64_t pcm_len, int n_windows, int win_len, int64_t step, const float* w,
                      const float* gamma, const float* beta, int channels, void* scale_shift, void* out,
                      int out_rows, void* stream);
/* The same with window w starting at sample win_offsets[w] ([dev] int64): windows of several files packed into one
 * call (the wav2vec2-family windows are independent of each other). */
SEGMA_API int segma_w2v2_layer0_at(const float* pcm, int64_t pcm_len, int n_windows, int win_len, const int64_t* win_offsets,
                         const float* w, const float* gamma, const float* beta, int channels, void* scale_shift,
                         void* out, int out_rows, void* stream);

/* WavLM gate on the relative-position bias (site-packages/torchaudio/models/wav2vec2/wavlm_attention.py:185-193):
 * x fp32 (rows = n_windows*T, n_heads*64) layer input; gate_w (8, 64), gate_b (8), gate_const (n_heads);
 * gate[(b*n_heads + h)*T + i] = ga*(gb*const_h - 1) + 2 -- the per-row factor segma_attention applies to pos_bias. */
SEGMA_API int segma_wavlm_gate(const float* x, int64_t rows, int T, int n_heads, const float* gate_w, const float* gate_b,
                     const float* gate_const, float* gate, void* stream);

/* ---- dense encoder building blocks ----------------------------------------------------------
 * Replace torch's dispatch of nn.Linear / nn.Conv1d / LayerNorm / SDPA inside
 * WhisperEncoder.forward (site-packages/transformers/models/whisper/modeling_whisper.py:593-647)
 * and torchaudio's wav2vec2 Encoder (site-packages/torchaudio/models/wav2vec2/components.py).
 */

/* C = epilogue(A * W^T): A fp16, W fp16 (n, k) row-major, fp32 accumulation on tcgen05 tensor cores (TMEM
 * accumulators, TMA-fed 128B-swizzled shared-memory stages).
 * Plain GEMM (conv_taps <= 1): A is (batch, rows_per_batch, k) with element strides (a_batch_stride,
 * a_row_stride, 1).
 * Implicit-GEMM Conv1d (conv_taps = kernel size, conv_stride = s): A is the time-major activation
 * (batch, a_rows_per_batch, c) with c = k / conv_taps channels, already padded by the caller; output row t of
 * a batch reads input rows s*t .. s*t + conv_taps - 1; W is (n, conv_taps * c) with the tap index major.
 * No im2col buffer is written: tap j is a TMA box at column (j % s) * c, row t + j / s of the activation
 * viewed s rows at a time.
 * epilogue: v = acc + bias[n]; if (flags & GELU) v = gelu_erf(v);
 *           if (add_src) v += add_src[(b*add_batch_rows + r) * n + col]   (fp32; add_batch_rows = 0: one
 *                        (rows_per_batch, n) table shared by every batch, e.g. position embeddings)
 *           out[(b*out_batch_rows + out_row_offset + r) * ldo + col] = v  (fp16, or fp32 with OUT_F32)
 */
#define SEGMA_GEMM_GELU 1
#define SEGMA_GEMM_OUT_F32 2
typedef struct {
  const void* a;          /* [dev] fp16, 16-byte aligned */
  int64_t a_batch_stride; /* elements, multiple of 8 */
  int64_t a_row_stride;   /* elements, multiple of 8 */
  int batch;
  int rows_per_batch;     /* output rows per batch */
  int a_rows_per_batch;   /* conv only: input rows per batch (multiple of conv_stride); 0 = rows_per_batch */
  int k;                  /* total reduction length (conv: conv_taps * channels) */
  int conv_taps;          /* 0 or 1: plain GEMM */
  int conv_stride;        /* 0 or 1: unit stride */
  const void* w;          /* [dev] fp16 (n, k) row-major */
  int n;                  /* multiple of 32 */
  const float* bias;      /* [dev] (n) or NULL */
  const float* add_src;   /* [dev] fp32 rows of n, or NULL; may alias out (residual update in place) */
  int64_t add_batch_rows; /* rows between consecutive batches in add_src */
  void* out;              /* [dev] 16-byte aligned */
  int64_t out_batch_rows;
  int64_t out_row_offset;
  int64_t ldo;            /* elements, multiple of 8 */
  int flags;
  int a_col_per_ntile;    /* grouped conv: extra A column offset per N tile (0 otherwise) */
  int a_cols;             /* columns of an activation row (0 = channels per tap); > channels for grouped conv */
  int force_bn;           /* 0 = auto; 128, 192 or 256 = N tile width */
} segma_gemm_args;
SEGMA_API int segma_gemm_f16(const segma_gemm_args* args, void* stream);

/* y = LayerNorm(x) * gamma + beta over the last dim (eps 1e-5), x fp32 (rows, d).
 *   out_f16 [dev] (rows, d) or NULL;  out_f32 [dev] (rows, d) fp32 or NULL (may alias x)
 *   mix      [dev] fp32 (rows/period, n_keep, d) or NULL: for rows r with (r % period) < n_keep,
 *            mix += w_in * x + w_out * y   (the layer-weighted sum of
 *            src/segma/models/whisper/surgical_hydra.py:82-98 restricted to the kept frames)
 *   mix_init: 1 = overwrite instead of accumulate
 *   only_kept: 1 = normalise only the rows with (r % period) < n_keep (the frames the heads read); the
 *              other rows of the outputs are left untouched
 */
SEGMA_API int segma_layernorm(const float* x, const float* gamma, const float* beta, int64_t rows, int d, void* out_f16,
                    float* out_f32, float* mix, int period, int n_keep, float w_in, float w_out, int mix_init,
                    int only_kept, void* stream);

/* softmax(Q K^T + bias) V per (window, head); qkv fp16 (n_windows*T, 3*d) rows = [q | k | v],
 * q pre-scaled, head_dim 64; out fp16 (n_windows*T, d).  Optional WavLM gated relative bias:
 * bias[b,h,i,j] = gate[(b*H + h)*T + i] * pos_bias[(h*T + i)*pos_bias_ld + j] (fp32), both NULL otherwise;
 * a row stride pos_bias_ld that is a multiple of 4 floats lets the tcgen05 kernel read it with 128-bit loads.
 * n_query: only the first n_query rows of each window are computed (<= T).
 */
SEGMA_API int segma_attention(const void* qkv, int n_windows, int T, int n_heads, int n_query, const float* gate,
                    const float* pos_bias, int pos_bias_ld, void* out, void* stream);

/* The same attention when the position table is Toeplitz, as WavLM's bucketed relative-position bias is
 * (site-packages/torchaudio/models/wav2vec2/wavlm_attention.py:85-139: the bucket depends on j - i only):
 * bias[b,h,i,j] = gate[(b*H + h)*T + i] * rel_bias[h*(2T-1) + (j - i + T - 1)].  The kernel keeps the slice of the
 * vector its 128 query rows need in shared memory instead of streaming T*T floats per head.  T <= 1024.
 */
SEGMA_API int segma_attention_rel(const void* qkv, int n_windows, int T, int n_heads, int n_query, const float* gate,
                        const float* rel_bias, void* out, void* stream);

/* fp32 -> fp16 copy of a (rows, cols) matrix with row strides. */
SEGMA_API int segma_cast_f16(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int cols, void* stream);

/* fp32 (rows, cols) -> fp16 (rows, 3*cols) = [hi | lo | hi], hi = fp16(x), lo = fp16(x - hi): the A operand of a
 * split-precision GEMM against a weight packed as [W_hi | W_hi | W_lo] (K = 3*cols).  Used for the LSTM input
 * projection (nn.LSTM's x W_ih^T, src/segma/models/whisper/surgical_hydra.py:57-60,101), whose rounding error is
 * otherwise a systematic perturbation that adds up along the 128-step recurrence. */
SEGMA_API int segma_cast_f16_split(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int cols, void* stream);

/* ---- LSTM over the window axis + per-label heads ---------------------------------------------
 * One direction-pair of one nn.LSTM layer (src/segma/models/whisper/hydra.py:48-51,81): the
 * sequence axis is the window batch (n_steps), the LSTM "batch" is the n_rows kept frames.
 *   pre   [dev] fp32 (n_steps, n_rows, n_dirs*4H): x W_ih^T + b_ih + b_hh, gate order i,f,g,o,
 *         forward direction first
 *   w_hh_t[dev] fp32 (n_dirs, H, 4H): W_hh transposed; kept at fp32 precision on the SM (H = 64: fp32 in shared
 *         memory; H = 128: fp16 hi in shared memory + fp16 lo in registers; H = 256: fp32 through L2)
 *   out   [dev] fp32 (n_steps, n_rows, n_dirs*H); out_f16 same shape or NULL
 */
SEGMA_API int segma_lstm_layer(const float* pre, const float* w_hh_t, int n_steps, int n_rows, int hidden, int n_dirs,
                     float* out, void* out_f16, void* stream);

/* logits[(frame_offset + s*step_frames + r) * C + c] = feat[s, r, :] . w[c, :] + b[c] for r < n_keep
 * (torch.stack of the per-label Linear(F, 1) heads, surgical_hydra.py:107-109), written straight
 * onto the file timeline when windows tile it, or per window (step_frames = n_keep) for stitching. */
SEGMA_API int segma_heads(const float* feat, int n_steps, int n_rows, int n_feat, int n_keep, const float* w, const float* b,
                int n_labels, float* logits, int64_t frame_offset, int step_frames, void* stream);
/* The same with window s writing frames frame_offsets[s] + r ([dev] int64): packed windows of several files. */
SEGMA_API int segma_heads_at(const float* feat, int n_steps, int n_rows, int n_feat, int n_keep, const float* w, const float* b,
                   int n_labels, float* logits, const int64_t* frame_offsets, void* stream);

/* ---- stitching and interval decoding --------------------------------------------------------
 * Overlap-add of per-window frame logits onto the file timeline: out[g] = mean over the windows
 * covering frame g, accumulated in window order in fp32; with step_frames == frames_per_window it is
 * the concatenation of src/segma/inference.py:209-211.  window_logits holds n_windows full windows
 * of frames_per_window frames followed by an optional tail window of tail_frames frames that starts
 * at frame n_windows*step_frames.
 */
SEGMA_API int segma_stitch(const float* window_logits, int n_windows, int frames_per_window, int step_frames,
                 int tail_frames, int n_labels, float* out, int64_t n_frames, void* stream);

/* Threshold + run-length decode of (n_frames, C) logits into an interval table, replacing
 * apply_thresholds + create_intervals (src/segma/inference.py:214-263):
 *   mode SEGMA_DECODE_SIGMOID: active = 1/(1+expf(-x)) > thresholds[c]   (fp32, strict)
 *   mode SEGMA_DECODE_LOGIT  : active = x > thresholds[c]                 (host-derived logit cut)
 * Frames of several files may be concatenated: file f owns frames [file_offsets[f], file_offsets[f+1]).
 * table rows are int32 (file, label, start_sample, end_sample), ordered by file, then label, then
 * time; start = 320*first_frame, end = 320*(last_frame+1), frames relative to the file.
 * *count [dev] receives the total number of intervals even if it exceeds `capacity` (rows beyond
 * capacity are not written).
 */
#define SEGMA_DECODE_SIGMOID 0
#define SEGMA_DECODE_LOGIT 1
SEGMA_API size_t segma_decode_workspace_bytes(int64_t n_frames, int n_files, int n_labels);
SEGMA_API int segma_decode_intervals(const float* logits, const int64_t* file_offsets /* [host] n_files+1 */, int n_files,
                           int n_labels, const float* thresholds /* [host] n_labels */, int mode, int32_t* table,
                           int64_t capacity, int32_t* count, void* workspace, size_t workspace_bytes, void* stream);
/* Extension (default off in the Python API): onset / offset hysteresis with the `upper_bound` that the reference
 * stores next to `lower_bound` but never reads (src/segma/inference.py:308-312).  A label turns on when its logit
 * exceeds onset_cuts[c], off when it is at or below offset_cuts[c], and otherwise keeps its state (off at the start
 * of every file).  Cuts are logit-domain [host]; onset >= offset; n_labels <= 8.  Same table / count / workspace
 * contract as segma_decode_intervals; with onset == offset the result is identical to it. */
SEGMA_API int segma_decode_intervals_hysteresis(const float* logits, const int64_t* file_offsets, int n_files, int n_labels,
                                      const float* offset_cuts, const float* onset_cuts, int32_t* table,
                                      int64_t capacity, int32_t* count, void* workspace, size_t workspace_bytes,
                                      void* stream);

/* Extension: interval-table post-processing on the device.  Rows of the same (file, label) separated by at most
 * max_gap_samples are merged (0 merges adjacent / overlapping rows, the semantics of the reference's unused
 * Intervals struct, src/segma/structs/interval.py:19-34), then rows shorter than min_duration_samples are dropped.
 * table [dev] (n, 4) sorted as segma_decode_intervals emits it; scratch, out [dev] (n, 4) / (capacity, 4);
 * counts [dev] int32[2] = {merged rows, rows written to out}. */
SEGMA_API int segma_postprocess_intervals(const int32_t* table, int64_t n, int max_gap_samples, int min_duration_samples,
                                int32_t* scratch, int32_t* out, int64_t capacity, int32_t* counts, void* stream);

/* The boolean mask of apply_thresholds alone: mask[f*C + c] (uint8). */
SEGMA_API int segma_threshold_mask(const float* logits, int64_t n_frames, int n_labels, const float* thresholds /* [host] */,
                         int mode, uint8_t* mask, void* stream);

/* ---- threshold tuning (scripts/tune.py:213-256) ------------------------------------------------------
 * One pass over (n_frames, C) logits and uint8 reference labels of the same shape: hist[c][y][b] counts the
 * frames of label c with reference y (0/1) whose logit exceeds exactly b of the n_cuts ascending logit-domain
 * cuts [host].  The prediction at grid point k is positive iff b > k, so TP/FP/FN of every threshold follow by
 * suffix sums.  hist [dev] (C, 2, n_cuts + 1) uint64, zeroed by the call.  n_cuts <= 128.
 */
SEGMA_API int segma_threshold_histogram(const float* logits, const uint8_t* truth, int64_t n_frames, int n_labels,
                              const float* cuts, int n_cuts, unsigned long long* hist, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEGMA_B200_H */
