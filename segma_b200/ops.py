"""Torch-tensor front of the C ABI: shape/dtype/device checks, raw pointers, the current CUDA stream.

PyTorch is only the allocator and stream provider here; all arithmetic happens in libsegma_b200.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native
from ._native import GemmArgs, SegmaNativeError, check

GEMM_GELU = 1
GEMM_OUT_F32 = 2
DECODE_SIGMOID = 0
DECODE_LOGIT = 1


def _lib():
    return _native.load()


class _Stats:
    """Launch accounting (bench.py's ``gpu_launches``) and optional CUDA-event timing of every C-ABI call."""

    def __init__(self):
        self.launches = 0
        self.profile = False
        self.events: list[tuple[str, torch.cuda.Event, torch.cuda.Event, float]] = []

    def reset(self):
        self.launches = 0
        self.events = []

    def breakdown(self) -> dict:
        """name -> {"ms": total, "calls": n, "work": summed flops/bytes}; call after a synchronize."""
        out: dict[str, dict] = {}
        for name, e0, e1, work in self.events:
            d = out.setdefault(name, {"ms": 0.0, "calls": 0, "work": 0.0})
            d["ms"] += e0.elapsed_time(e1)
            d["calls"] += 1
            d["work"] += work
        return out


stats = _Stats()


def _call(name: str, n_launches: int, fn, *args, work: float = 0.0) -> None:
    if stats.profile:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        stats.events.append((name, e0, e1, work))
    else:
        rc = fn(*args)
    check(rc, name.split("[")[0])
    stats.launches += n_launches


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _dev(t: torch.Tensor, dtype, name: str) -> int:
    if not t.is_cuda:
        raise SegmaNativeError(f"{name} must be a CUDA tensor (segma_b200 has no CPU path)")
    if t.device.index != torch.cuda.current_device():
        # kernels launch on the current device's stream; engines and the inference entry points select the tensors'
        # device themselves (torch.cuda.device), direct callers of ops must do the same
        raise SegmaNativeError(f"{name} lives on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}; "
                               f"wrap the call in torch.cuda.device({t.device.index})")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    return t.data_ptr()


def _ptr(t, dtype, name):
    return None if t is None else _dev(t, dtype, name)


def device_check() -> None:
    check(_lib().segma_device_check(), "segma_device_check")


# ---- log-mel -------------------------------------------------------------------------------------
def logmel(pcm: torch.Tensor, n_windows: int, win_len: int, step: int, out_f32: bool = True, out_tm: bool = False,
           pcm_offset: int = 0):
    """Windows ``pcm[pcm_offset + i*step : ... + win_len]`` -> Whisper log-mel.
    Returns ``(f32 (n,80,3000) | None, fp16 time-major (n,3002,80) | None)``."""
    lib = _lib()
    _dev(pcm, torch.float32, "pcm")
    assert pcm.dim() == 1 and pcm.is_contiguous()
    f32 = torch.empty((n_windows, 80, 3000), dtype=torch.float32, device=pcm.device) if out_f32 else None
    tm = torch.empty((n_windows, 3002, 80), dtype=torch.float16, device=pcm.device) if out_tm else None
    nbytes = lib.segma_logmel_scratch_bytes(n_windows, win_len)
    scratch = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=pcm.device)
    view = pcm[pcm_offset:]
    _call("segma_logmel", 3, _lib().segma_logmel, view.data_ptr(), view.numel(), n_windows, win_len, step,
                         None if f32 is None else f32.data_ptr(), None if tm is None else tm.data_ptr(),
                         scratch.data_ptr(), _stream())
    return f32, tm


def logmel_into(pcm_view: torch.Tensor, n_windows: int, win_len: int, step: int, tm: torch.Tensor,
                scratch: torch.Tensor) -> None:
    _call("segma_logmel", 3, _lib().segma_logmel, pcm_view.data_ptr(), pcm_view.numel(), n_windows, win_len, step, None, tm.data_ptr(),
                            scratch.data_ptr(), _stream())


def logmel_scratch_bytes(n_windows: int, win_len: int) -> int:
    return _lib().segma_logmel_scratch_bytes(n_windows, win_len)


def mel_filters() -> np.ndarray:
    out = np.empty((201, 80), dtype=np.float32)
    check(_lib().segma_logmel_get_filters(out.ctypes.data), "segma_logmel_get_filters")
    return out


def set_mel_filters(mel: np.ndarray) -> None:
    m = np.ascontiguousarray(mel, dtype=np.float32)
    assert m.shape == (201, 80)
    check(_lib().segma_logmel_set_filters(m.ctypes.data), "segma_logmel_set_filters")


# ---- GEMM / conv ---------------------------------------------------------------------------------
def gemm_raw(a_ptr, a_batch_stride, a_row_stride, batch, rows_per_batch, k, w, n, out_ptr, ldo, *, bias=None,
             add_src_ptr=None, add_batch_rows=0, out_batch_rows=None, out_row_offset=0, flags=0, conv_taps=0,
             conv_stride=0, a_rows_per_batch=0, a_col_per_ntile=0, a_cols=0, force_bn=0, work=None) -> None:
    """``work``: algorithmic FLOPs of the launch for the roofline accounting when they differ from 2*M*N*K of the
    launched shape (a grouped convolution packed with zero blocks executes more than it computes)."""
    args = GemmArgs(
        a=a_ptr, a_batch_stride=a_batch_stride, a_row_stride=a_row_stride, batch=batch,
        rows_per_batch=rows_per_batch, a_rows_per_batch=a_rows_per_batch, k=k, conv_taps=conv_taps,
        conv_stride=conv_stride, w=_dev(w, torch.float16, "w"), n=n, bias=_ptr(bias, torch.float32, "bias"),
        add_src=add_src_ptr, add_batch_rows=add_batch_rows, out=out_ptr,
        out_batch_rows=rows_per_batch if out_batch_rows is None else out_batch_rows,
        out_row_offset=out_row_offset, ldo=ldo, flags=flags, a_col_per_ntile=a_col_per_ntile, a_cols=a_cols, force_bn=force_bn,
    )
    name = "segma_gemm_f16"
    if stats.profile:
        name += f"[n{n} k{k}{' conv' if conv_taps > 1 else ''}{' gelu' if flags & GEMM_GELU else ''}" \
                f"{' +src' if add_src_ptr else ''}{' f32' if flags & GEMM_OUT_F32 else ''}]"
    _call(name, 1, _lib().segma_gemm_f16, C.byref(args), _stream(),
          work=2.0 * batch * rows_per_batch * n * k if work is None else work)


def linear(a: torch.Tensor, w: torch.Tensor, bias=None, *, gelu=False, add_src=None, out=None, out_f32=False,
           force_bn=0) -> torch.Tensor:
    """out = epilogue(a @ w.T); a (M, K) fp16, w (N, K) fp16."""
    _dev(a, torch.float16, "a")
    assert a.dim() == 2 and a.stride(1) == 1 and w.is_contiguous()
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32 if out_f32 else torch.float16, device=a.device)
    out_f32 = out.dtype == torch.float32
    flags = (GEMM_GELU if gelu else 0) | (GEMM_OUT_F32 if out_f32 else 0)
    gemm_raw(a.data_ptr(), 0, a.stride(0), 1, M, K, w, N, out.data_ptr(), out.stride(0), bias=bias,
             add_src_ptr=None if add_src is None else _dev(add_src, torch.float32, "add_src"), flags=flags,
             force_bn=force_bn)
    return out


def linear_rows(a: torch.Tensor, n_batch: int, batch_rows: int, keep: int, w: torch.Tensor, bias, out: torch.Tensor, *,
                gelu=False, add_src=None) -> None:
    """``linear`` on the first ``keep`` rows of every ``batch_rows``-row block of a (n_batch*batch_rows, K);
    out / add_src use the same row blocking (rows >= keep of each block are left untouched)."""
    K, N = a.shape[1], w.shape[0]
    flags = (GEMM_GELU if gelu else 0) | (GEMM_OUT_F32 if out.dtype == torch.float32 else 0)
    gemm_raw(_dev(a, torch.float16, "a"), batch_rows * a.stride(0), a.stride(0), n_batch, keep, K, w, N, out.data_ptr(),
             out.stride(0), bias=bias, add_src_ptr=None if add_src is None else _dev(add_src, torch.float32, "add_src"),
             add_batch_rows=batch_rows, out_batch_rows=batch_rows, flags=flags)


def conv1d_tm(x_tm: torch.Tensor, w_tap_major: torch.Tensor, bias, taps: int, stride: int, out_rows: int, *,
              gelu=True, add_src=None, add_batch_rows=0, out=None, out_batch_rows=None, out_row_offset=0, out_f32=False):
    """Implicit-GEMM Conv1d on a padded time-major activation x_tm (B, rows_in, C) fp16;
    w_tap_major (N, taps*C) fp16.  Output (B, out_batch_rows, N) rows [out_row_offset, +out_rows)."""
    B, rows_in, Cc = x_tm.shape
    assert x_tm.is_contiguous()
    N = w_tap_major.shape[0]
    obr = out_rows if out_batch_rows is None else out_batch_rows
    if out is None:
        out = torch.zeros((B, obr, N), dtype=torch.float32 if out_f32 else torch.float16, device=x_tm.device)
    flags = (GEMM_GELU if gelu else 0) | (GEMM_OUT_F32 if out.dtype == torch.float32 else 0)
    gemm_raw(_dev(x_tm, torch.float16, "x_tm"), rows_in * Cc, Cc, B, out_rows, taps * Cc, w_tap_major, N,
             out.data_ptr(), N, bias=bias,
             add_src_ptr=None if add_src is None else _dev(add_src, torch.float32, "add_src"), add_batch_rows=add_batch_rows,
             out_batch_rows=obr, out_row_offset=out_row_offset, flags=flags, conv_taps=taps, conv_stride=stride,
             a_rows_per_batch=rows_in)
    return out


# ---- layernorm / cast / attention ----------------------------------------------------------------
def layernorm(x: torch.Tensor, gamma, beta, *, out_f16=None, out_f32=None, mix=None, period=1, n_keep=0, w_in=0.0,
              w_out=0.0, mix_init=False, only_kept=False) -> None:
    rows, d = x.shape
    assert x.is_contiguous()
    _call("segma_layernorm", 1, _lib().segma_layernorm, _dev(x, torch.float32, "x"), _dev(gamma, torch.float32, "gamma"),
                               _dev(beta, torch.float32, "beta"), rows, d, _ptr(out_f16, torch.float16, "out_f16"),
                               _ptr(out_f32, torch.float32, "out_f32"), _ptr(mix, torch.float32, "mix"), period,
                               n_keep, float(w_in), float(w_out), int(mix_init), int(only_kept), _stream())


def cast_f16(src: torch.Tensor, dst: torch.Tensor) -> None:
    rows, cols = src.shape
    _call("segma_cast_f16", 1, _lib().segma_cast_f16, _dev(src, torch.float32, "src"), src.stride(0), _dev(dst, torch.float16, "dst"),
                               dst.stride(0), rows, cols, _stream())


def cast_f16_split(src: torch.Tensor, dst: torch.Tensor) -> None:
    """(rows, cols) fp32 -> (rows, 3*cols) fp16 ``[hi | lo | hi]`` (see ``split_weight``)."""
    rows, cols = src.shape
    assert dst.shape == (rows, 3 * cols)
    _call("segma_cast_f16_split", 1, _lib().segma_cast_f16_split, _dev(src, torch.float32, "src"), src.stride(0),
          _dev(dst, torch.float16, "dst"), dst.stride(0), rows, cols, _stream())


def split_weight(w: torch.Tensor) -> torch.Tensor:
    """(N, K) fp32 weight -> (N, 3K) fp16 ``[W_hi | W_hi | W_lo]`` matching ``cast_f16_split``'s ``[hi | lo | hi]``."""
    w = w.detach().float()
    hi = w.to(torch.float16)
    lo = (w - hi.float()).to(torch.float16)
    return torch.cat([hi, hi, lo], dim=1).contiguous()


def attention(qkv: torch.Tensor, n_windows: int, T: int, n_heads: int, *, n_query=None, gate=None, pos_bias=None,
              rel_bias=None, out=None) -> torch.Tensor:
    """``pos_bias`` (H, T, T) or, for a Toeplitz table, ``rel_bias`` (H, 2T-1) with bias[i, j] = rel[j - i + T - 1]."""
    assert qkv.is_contiguous() and qkv.shape == (n_windows * T, 3 * n_heads * 64)
    if out is None:
        out = torch.zeros((n_windows * T, n_heads * 64), dtype=torch.float16, device=qkv.device)
    if rel_bias is not None:
        assert pos_bias is None and gate is not None
        assert rel_bias.is_contiguous() and rel_bias.shape == (n_heads, 2 * T - 1)
        _call("segma_attention_rel", 1, _lib().segma_attention_rel, _dev(qkv, torch.float16, "qkv"), n_windows, T, n_heads,
              T if n_query is None else n_query, _ptr(gate, torch.float32, "gate"),
              _ptr(rel_bias, torch.float32, "rel_bias"), _dev(out, torch.float16, "out"), _stream(),
              work=4.0 * n_windows * n_heads * (T if n_query is None else n_query) * T * 64)
        return out
    pb_ld = 0
    if pos_bias is not None:
        assert pos_bias.dim() == 3 and pos_bias.shape[:2] == (n_heads, T) and pos_bias.stride(2) == 1
        assert pos_bias.stride(0) == T * pos_bias.stride(1)
        pb_ld = pos_bias.stride(1)
    _call("segma_attention", 1, _lib().segma_attention, _dev(qkv, torch.float16, "qkv"), n_windows, T, n_heads,
          T if n_query is None else n_query, _ptr(gate, torch.float32, "gate"), _ptr(pos_bias, torch.float32, "pos_bias"),
          pb_ld, _dev(out, torch.float16, "out"), _stream(),
          work=4.0 * n_windows * n_heads * (T if n_query is None else n_query) * T * 64)
    return out


# ---- LSTM / heads ------------------------------------------------------------------------------
def lstm_layer(pre: torch.Tensor, w_hh_t: torch.Tensor, hidden: int, *, out=None, out_f16=None) -> torch.Tensor:
    n_steps, n_rows, g = pre.shape
    n_dirs = g // (4 * hidden)
    assert pre.is_contiguous() and w_hh_t.is_contiguous() and w_hh_t.shape == (n_dirs, hidden, 4 * hidden)
    if out is None:
        out = torch.empty((n_steps, n_rows, n_dirs * hidden), dtype=torch.float32, device=pre.device)
    _call("segma_lstm_layer", 1, _lib().segma_lstm_layer, _dev(pre, torch.float32, "pre"), _dev(w_hh_t, torch.float32, "w_hh_t"), n_steps, n_rows,
                                hidden, n_dirs, _dev(out, torch.float32, "out"),
                                _ptr(out_f16, torch.float16, "out_f16"), _stream())
    return out


def heads(feat: torch.Tensor, w: torch.Tensor, b: torch.Tensor, logits: torch.Tensor, frame_offset: int,
          step_frames: int, n_keep: int, frame_offsets: torch.Tensor | None = None) -> None:
    n_steps, n_rows, n_feat = feat.shape
    assert feat.is_contiguous() and w.is_contiguous() and logits.is_contiguous()
    if frame_offsets is not None:  # window s writes frames frame_offsets[s] + r (packed windows of several files)
        assert frame_offsets.numel() >= n_steps and frame_offsets.is_contiguous()
        _call("segma_heads", 1, _lib().segma_heads_at, _dev(feat, torch.float32, "feat"), n_steps, n_rows, n_feat, n_keep,
              _dev(w, torch.float32, "w"), _dev(b, torch.float32, "b"), w.shape[0], _dev(logits, torch.float32, "logits"),
              _dev(frame_offsets, torch.int64, "frame_offsets"), _stream())
        return
    _call("segma_heads", 1, _lib().segma_heads, _dev(feat, torch.float32, "feat"), n_steps, n_rows, n_feat, n_keep,
                           _dev(w, torch.float32, "w"), _dev(b, torch.float32, "b"), w.shape[0],
                           _dev(logits, torch.float32, "logits"), frame_offset, step_frames, _stream())


# ---- stitch / decode -----------------------------------------------------------------------------
def stitch(window_logits: torch.Tensor, n_windows: int, frames_per_window: int, step_frames: int, tail_frames: int,
           n_frames: int) -> torch.Tensor:
    C_ = window_logits.shape[-1]
    assert window_logits.is_contiguous()
    out = torch.empty((n_frames, C_), dtype=torch.float32, device=window_logits.device)
    _call("segma_stitch", 1, _lib().segma_stitch, _dev(window_logits, torch.float32, "window_logits"), n_windows, frames_per_window,
                            step_frames, tail_frames, C_, out.data_ptr(), n_frames, _stream())
    return out


def threshold_mask(logits: torch.Tensor, thresholds, mode: int = DECODE_SIGMOID) -> torch.Tensor:
    n, C_ = logits.shape
    assert logits.is_contiguous()
    thr = (C.c_float * C_)(*[float(t) for t in thresholds])
    mask = torch.empty((n, C_), dtype=torch.uint8, device=logits.device)
    _call("segma_threshold_mask", 1, _lib().segma_threshold_mask, _dev(logits, torch.float32, "logits"), n, C_, thr, mode, mask.data_ptr(), _stream())
    return mask.bool()


def decode_intervals(logits: torch.Tensor, thresholds, *, file_offsets=None, mode: int = DECODE_SIGMOID,
                     capacity: int | None = None, onset=None) -> torch.Tensor:
    """(n_frames, C) logits -> int32 (n_intervals, 4) table (file, label, start_sample, end_sample) on the device.
    Reads the interval count back once (the only synchronisation); retries if ``capacity`` was too small.
    ``onset`` (logit-domain cuts >= ``thresholds``) switches to onset / offset hysteresis."""
    lib = _lib()
    n, C_ = logits.shape
    assert logits.is_contiguous()
    offs = [0, n] if file_offsets is None else [int(v) for v in file_offsets]
    n_files = len(offs) - 1
    assert offs[-1] == n
    off_arr = (C.c_int64 * len(offs))(*offs)
    thr = (C.c_float * C_)(*[float(t) for t in thresholds])
    ws_bytes = lib.segma_decode_workspace_bytes(n, n_files, C_)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=logits.device)
    count = torch.zeros(1, dtype=torch.int32, device=logits.device)
    cap = max(1024, n // 16) if capacity is None else capacity
    while True:
        table = torch.empty((cap, 4), dtype=torch.int32, device=logits.device)
        if onset is None:
            _call("segma_decode_intervals", 3, lib.segma_decode_intervals, _dev(logits, torch.float32, "logits"), off_arr,
                  n_files, C_, thr, mode, table.data_ptr(), cap, count.data_ptr(), ws.data_ptr(), ws_bytes, _stream())
        else:
            hi = (C.c_float * C_)(*[float(t) for t in onset])
            _call("segma_decode_intervals_hysteresis", 6, lib.segma_decode_intervals_hysteresis,
                  _dev(logits, torch.float32, "logits"), off_arr, n_files, C_, thr, hi, table.data_ptr(), cap,
                  count.data_ptr(), ws.data_ptr(), ws_bytes, _stream())
        total = int(count.item())
        if total <= cap:
            return table[:total]
        cap = total


def decode_intervals_async(logits: torch.Tensor, thresholds, *, file_offsets=None, mode: int = DECODE_SIGMOID,
                           onset=None) -> tuple[torch.Tensor, torch.Tensor]:
    """``decode_intervals`` without the host read-back: returns ``(table, count)`` where ``table`` has room for the
    worst case (alternating frames: ceil(n_f / 2) runs per label and file) and ``count`` is a 1-element int32 device
    tensor; rows ``[count:]`` are undefined.  Nothing synchronises, so files can be queued back to back."""
    lib = _lib()
    n, C_ = logits.shape
    assert logits.is_contiguous()
    offs = [0, n] if file_offsets is None else [int(v) for v in file_offsets]
    n_files = len(offs) - 1
    assert offs[-1] == n
    off_arr = (C.c_int64 * len(offs))(*offs)
    thr = (C.c_float * C_)(*[float(t) for t in thresholds])
    ws_bytes = lib.segma_decode_workspace_bytes(n, n_files, C_)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=logits.device)
    count = torch.zeros(1, dtype=torch.int32, device=logits.device)
    cap = max(1, C_ * (n // 2 + n_files))
    table = torch.empty((cap, 4), dtype=torch.int32, device=logits.device)
    if n == 0:
        return table, count
    if onset is None:
        _call("segma_decode_intervals", 3, lib.segma_decode_intervals, _dev(logits, torch.float32, "logits"), off_arr,
              n_files, C_, thr, mode, table.data_ptr(), cap, count.data_ptr(), ws.data_ptr(), ws_bytes, _stream())
    else:
        hi = (C.c_float * C_)(*[float(t) for t in onset])
        _call("segma_decode_intervals_hysteresis", 6, lib.segma_decode_intervals_hysteresis,
              _dev(logits, torch.float32, "logits"), off_arr, n_files, C_, thr, hi, table.data_ptr(), cap,
              count.data_ptr(), ws.data_ptr(), ws_bytes, _stream())
    return table, count


# ---- wav2vec2 / WavLM ------------------------------------------------------------------------------
def w2v2_layer0(pcm_view: torch.Tensor, n_windows: int, win_len: int, step: int, w: torch.Tensor, gamma, beta,
                scale_shift: torch.Tensor, out: torch.Tensor, win_offsets: torch.Tensor | None = None) -> None:
    """out (n_windows, out_rows, C) fp16 = gelu(GroupNorm(conv_k10_s5(window))); window i starts at sample
    ``i * step`` of ``pcm_view``, or at ``win_offsets[i]`` (int64 device tensor) for packed windows of several files."""
    C_ = w.shape[0]
    assert out.is_contiguous() and out.shape[0] >= n_windows and out.shape[2] == C_
    rows0 = (win_len - 10) // 5 + 1
    work = 2.0 * n_windows * rows0 * C_ * 10
    if win_offsets is None:
        _call("segma_w2v2_layer0", 2, _lib().segma_w2v2_layer0, _dev(pcm_view, torch.float32, "pcm"), pcm_view.numel(),
              n_windows, win_len, step, _dev(w, torch.float32, "w"), _dev(gamma, torch.float32, "gamma"),
              _dev(beta, torch.float32, "beta"), C_, scale_shift.data_ptr(), _dev(out, torch.float16, "out"), out.shape[1],
              _stream(), work=work)
    else:
        assert win_offsets.numel() >= n_windows and win_offsets.is_contiguous()
        _call("segma_w2v2_layer0", 2, _lib().segma_w2v2_layer0_at, _dev(pcm_view, torch.float32, "pcm"), pcm_view.numel(),
              n_windows, win_len, _dev(win_offsets, torch.int64, "win_offsets"), _dev(w, torch.float32, "w"),
              _dev(gamma, torch.float32, "gamma"), _dev(beta, torch.float32, "beta"), C_, scale_shift.data_ptr(),
              _dev(out, torch.float16, "out"), out.shape[1], _stream(), work=work)


def wavlm_gate(x: torch.Tensor, T: int, n_heads: int, gate_w, gate_b, gate_const, gate: torch.Tensor) -> None:
    _call("segma_wavlm_gate", 1, _lib().segma_wavlm_gate, _dev(x, torch.float32, "x"), x.shape[0], T, n_heads,
          _dev(gate_w, torch.float32, "gate_w"), _dev(gate_b, torch.float32, "gate_b"),
          _dev(gate_const, torch.float32, "gate_const"), _dev(gate, torch.float32, "gate"), _stream())


# ---- audio staging -----------------------------------------------------------------------------------
PCM_S16, PCM_S32, PCM_F32 = 0, 1, 2


def pcm_to_f32(raw: torch.Tensor, fmt: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """Device tensor of native-width samples (int16 / int32 / float32) -> float32 in [-1, 1)."""
    want = {PCM_S16: torch.int16, PCM_S32: torch.int32, PCM_F32: torch.float32}[fmt]
    if out is None:
        out = torch.empty(raw.numel(), dtype=torch.float32, device=raw.device)
    _call("segma_pcm_to_f32", 1, _lib().segma_pcm_to_f32, _dev(raw, want, "raw"), fmt, raw.numel(),
          _dev(out, torch.float32, "out"), _stream())
    return out


# ---- threshold tuning -------------------------------------------------------------------------------
def threshold_histogram(logits: torch.Tensor, truth: torch.Tensor, cuts) -> torch.Tensor:
    """(n, C) logits + (n, C) uint8 reference labels -> int64 (C, 2, K+1) histogram of "cuts exceeded"."""
    n, C_ = logits.shape
    assert logits.is_contiguous() and truth.is_contiguous() and truth.shape == logits.shape
    K = len(cuts)
    arr = (C.c_float * K)(*[float(c) for c in cuts])
    hist = torch.empty((C_, 2, K + 1), dtype=torch.int64, device=logits.device)
    _call("segma_threshold_histogram", 1, _lib().segma_threshold_histogram, _dev(logits, torch.float32, "logits"),
          _dev(truth, torch.uint8, "truth"), n, C_, arr, K, hist.data_ptr(), _stream())
    return hist


def postprocess_intervals(table: torch.Tensor, max_gap_samples: int = 0, min_duration_samples: int = 0) -> torch.Tensor:
    """Merge rows of the same (file, label) closer than ``max_gap_samples`` and drop rows shorter than
    ``min_duration_samples``; (n, 4) int32 device table in the decode order -> filtered table."""
    n = table.shape[0]
    if n == 0:
        return table
    assert table.is_contiguous() and table.dtype == torch.int32
    scratch = torch.empty_like(table)
    out = torch.empty_like(table)
    counts = torch.zeros(2, dtype=torch.int32, device=table.device)
    _call("segma_postprocess_intervals", 2, _lib().segma_postprocess_intervals, table.data_ptr(), n,
          int(max_gap_samples), int(min_duration_samples), scratch.data_ptr(), out.data_ptr(), n, counts.data_ptr(),
          _stream())
    return out[: int(counts[1].item())]
