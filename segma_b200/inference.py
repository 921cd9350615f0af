"""Sliding-window inference driver with the reference's prediction API.

Same entry points and signatures as /root/reference/src/segma/inference.py -- ``apply_model_on_audio``
(119-211), ``apply_thresholds`` (214-234), ``create_intervals`` (237-263), ``write_intervals`` (266-283),
``infer_file`` (286-357), ``get_list_of_files_to_process`` (360-395), ``run_inference_on_audios`` (398-459)
and the ``python -m`` CLI (462-501) -- but the whole path from PCM to the interval table runs in
libsegma_b200 on the GPU: the file is staged to HBM once, windows are cut by the front-end kernel's
addressing (no ``unfold`` copy), frame logits are written straight onto the file timeline, and thresholds
+ run-length decoding happen on the device; only the (small) interval table crosses back.

Extensions over the reference (all default to its behaviour): ``window_step`` for overlapping windows
(logit-domain mean), in-memory audio instead of a path, and batched multi-file decoding.
"""
from __future__ import annotations

import argparse
import os
from logging import Logger
from pathlib import Path
from typing import Literal

import numpy as np
import torch
import yaml

from . import ops
from .annotation import rttm_line
from .config import Config, load_config
from .encoders import MultiLabelEncoder
from .geometry import (FRAME_SAMPLES, INFERENCE_SETTINGS, Chunkyfier, ConvolutionSettings, assign_units, batch_frame_range,
                       conv_frames, plan_packed_calls, plan_windows, plan_work_units)
from .io import PcmSource, get_audio_info, get_samples_in_range, stage_to_device
from .engine import resolve_device
from .models import BaseSegmentationModel, Models
from .thresholds import logit_cut

__all__ = [
    "Chunkyfier", "prepare_audio", "apply_model_on_audio", "apply_model_on_audios", "apply_thresholds", "create_intervals",
    "decode_logits", "write_intervals", "infer_file", "infer_corpus", "get_list_of_files_to_process",
    "run_inference_on_audios",
]


#: concurrent batches per file (streams / workspace slots).  Measured on B200: the path runs at the 1 kW power cap,
#: so overlapping batches buys nothing (2.26 vs 2.22 audio-h/s); the default stays strictly serial.
N_STREAMS = int(os.environ.get("SEGMA_STREAMS", "1"))
_STREAMS: dict[tuple, list] = {}
#: files queued on the device before the host waits for the oldest one's interval table
MAX_FILES_IN_FLIGHT = 4
#: ``infer_corpus``: streams that short files (fewer than SMALL_FILE_WINDOWS windows) of a Whisper-family corpus take
#: turns on; their forward calls are too small to fill the GPU one at a time
FILE_STREAMS = int(os.environ.get("SEGMA_FILE_STREAMS", "4"))
SMALL_FILE_WINDOWS = 32


def _side_streams(dev: torch.device, n: int):
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), n)
    if key not in _STREAMS:
        _STREAMS[key] = [torch.cuda.Stream(device=dev) for _ in range(n)]
    return _STREAMS[key]


def _cuda_device(device) -> torch.device:
    dev = torch.device("cuda" if device in ("gpu", None) else device)
    if dev.type != "cuda":
        raise ops.SegmaNativeError(f"device '{device}': segma_b200 runs on CUDA (sm_100a) only; there is no CPU path")
    return resolve_device(dev)


def prepare_audio(audio_path, model: BaseSegmentationModel, device, start_f: int, end_f: int | None = None):
    """Samples ``[start_f, end_f)`` of the file as a 1-D fp32 device tensor (inference.py:92-116).
    The Whisper hook is *not* applied here: the reference applies it to the whole span, which cannot
    work (SURVEY.md finding 5); the log-mel is computed per window by the front-end kernel instead."""
    num = end_f - start_f if end_f else -1
    t = get_samples_in_range(audio_path, start_f=start_f, duration_f=num)
    if t.shape[0] != 1:
        raise ValueError(f"only mono audio is supported, got {t.shape[0]} channels")
    host = t.reshape(-1)
    if torch.cuda.is_available():
        host = host.pin_memory()
    return host.to(_cuda_device(device), non_blocking=True)


def apply_model_on_audio(
    audio_path,
    model: BaseSegmentationModel,
    conv_settings: ConvolutionSettings,
    device: Literal["cuda", "gpu"] = "cuda",
    batch_size: int = 128,
    chunk_duration_s: float = 4.0,
    sample_rate: int = 16_000,
    window_step: int | None = None,
    slot_base: int = 0,
    batch_range: tuple[int, int] | None = None,
) -> torch.Tensor:
    """Apply model on audio, return a ``(n_frames, n_classes)`` fp32 tensor of raw logits on the device.

    Windows and batches are exactly the reference's (inference.py:129-206): ``batch_size`` consecutive
    windows per forward call, one remainder batch, then the tail alone -- the LSTM of the Whisper-family
    models couples the windows of a call, so batch boundaries are part of the result (SURVEY.md finding 6).
    ``audio_path`` may also be a 1-D float32 array / tensor (host or device).  ``batch_range=(lo, hi)`` runs only
    forward calls lo ... hi-1 of the file and returns the logits of the frames they cover
    (``geometry.batch_frame_range``): a rank's share of a very long file.
    """
    dev = _cuda_device(device)
    engine = model._require_engine()
    if engine.device != dev:
        raise ops.SegmaNativeError(f"model weights are on {engine.device} but device={dev} was requested; call model.to(device)")
    with torch.cuda.device(dev):
        return _apply_model_on_audio(audio_path, model, engine, conv_settings, dev, batch_size, chunk_duration_s,
                                     sample_rate, window_step, slot_base, batch_range)


def _apply_model_on_audio(audio_path, model, engine, conv_settings, dev, batch_size, chunk_duration_s, sample_rate,
                          window_step, slot_base=0, batch_range=None) -> torch.Tensor:
    chunk_f = int(chunk_duration_s * sample_rate)
    chunky = Chunkyfier(batch_size, chunk_f, conv_settings)  # same derived quantities as the reference
    step = chunky.step if window_step is None else int(window_step)
    # native-width PCM over PCIe, widened on the device, staged batch by batch behind the compute of earlier batches
    source = PcmSource(audio_path, dev)
    pcm = source.dev
    n_samples = source.n_samples
    frames_per_window = model.n_keep if model.family == "whisper" else conv_frames(chunk_f)
    plan = plan_windows(n_samples, chunk_f, batch_size, step, frames_per_window)
    n_labels = model.label_encoder.n_labels
    sf = plan.step_frames
    tiled = sf == frames_per_window  # windows tile the frame grid: stitching is concatenation
    if plan.n_frames == 0:
        return torch.zeros((0, n_labels), dtype=torch.float32, device=dev)
    batches, f_lo = plan.batches, 0
    if batch_range is not None:
        if not tiled:
            raise ValueError("batch_range needs windows that tile the frame grid (no window_step overlap)")
        batches = plan.batches[batch_range[0]: batch_range[1]]
        f_lo, f_hi = batch_frame_range(plan, *batch_range)
        if not batches:
            return torch.zeros((0, n_labels), dtype=torch.float32, device=dev)
        source.skip_to(batches[0].start_sample)
    if tiled:
        logits = torch.empty((plan.n_frames if batch_range is None else f_hi - f_lo, n_labels), dtype=torch.float32, device=dev)
    else:
        n_full = sum(b.n_windows for b in plan.batches if not b.is_tail)
        tail = next((b for b in plan.batches if b.is_tail), None)
        win_logits = torch.empty((n_full * frames_per_window + (tail.frames_per_window if tail else 0), n_labels),
                                 dtype=torch.float32, device=dev)
    # Batches are independent of one another (the LSTM couples windows *inside* a batch only), so consecutive
    # batches alternate between two streams / workspace slots: one batch's HBM- and latency-bound kernels
    # (LayerNorm, LSTM, heads) and kernel tails overlap the other's tensor-core kernels.
    main = torch.cuda.current_stream(dev)
    n_lanes = max(1, min(N_STREAMS, len(batches)))
    lanes = _side_streams(dev, n_lanes) if n_lanes > 1 else [main]
    for st in lanes:
        if st is not main:
            st.wait_stream(main)
    target = logits if tiled else win_logits
    for i, b in enumerate(batches):
        source.ensure(b.start_sample + (b.n_windows - 1) * step + b.win_len)
        if lanes[i % n_lanes] is not main:
            lanes[i % n_lanes].wait_stream(main)
        with torch.cuda.stream(lanes[i % n_lanes]):
            if tiled:
                engine.forward_pcm(pcm, b.start_sample, b.n_windows, b.win_len, step, logits, b.first_window * sf - f_lo, sf,
                                   b.frames_per_window, slot=slot_base + i % n_lanes)
            else:
                engine.forward_pcm(pcm, b.start_sample, b.n_windows, b.win_len, step, win_logits,
                                   b.first_window * frames_per_window, frames_per_window, b.frames_per_window,
                                   slot=slot_base + i % n_lanes)
    for st in lanes:
        if st is not main:
            main.wait_stream(st)
            target.record_stream(st)
            pcm.record_stream(st)
    if tiled:
        return logits
    n_full = sum(b.n_windows for b in plan.batches if not b.is_tail)
    tail_frames = next((b.frames_per_window for b in plan.batches if b.is_tail), 0)
    return ops.stitch(win_logits, n_full, frames_per_window, sf, tail_frames, plan.n_frames)


#: windows per packed group of files (a few forward calls of ``batch_size`` windows) and its PCM budget
PACK_MAX_WINDOWS = 1024


def apply_model_on_audios(audios, model: BaseSegmentationModel, conv_settings: ConvolutionSettings = INFERENCE_SETTINGS,
                          device: Literal["cuda", "gpu"] = "cuda", batch_size: int = 128, chunk_duration_s: float = 4.0,
                          sample_rate: int = 16_000) -> list[torch.Tensor]:
    """``apply_model_on_audio`` for several files at once: one ``(n_frames_f, n_classes)`` logits tensor per file.

    For the wav2vec2 family (HuBERT / WavLM: no recurrence over the window axis, so a window's logits do not depend
    on what else is in its forward call) the windows of all files are packed into forward calls of ``batch_size``
    windows regardless of file boundaries, and tails of equal length share a call: a corpus of short clips runs at the
    batch efficiency of long files instead of one partial batch + one tail per file (the reference loops over files,
    inference.py:442-458, and over a file's batches, 138-206).  Results are the same bits as file-by-file calls.
    Whisper-family models couple the windows of a call through the LSTM (SURVEY.md finding 6): their batch boundaries
    are part of the result, so files are processed one after the other."""
    dev = _cuda_device(device)
    if model.family != "wav2vec2":
        return [apply_model_on_audio(a, model, conv_settings, dev, batch_size, chunk_duration_s, sample_rate) for a in audios]
    engine = model._require_engine()
    if engine.device != dev:
        raise ops.SegmaNativeError(f"model weights are on {engine.device} but device={dev} was requested; call model.to(device)")
    from .io import audio_n_samples

    chunk_f = int(chunk_duration_s * sample_rate)
    step = Chunkyfier(batch_size, chunk_f, conv_settings).step
    fpw = conv_frames(chunk_f)
    n_labels = model.label_encoder.n_labels
    out: list[torch.Tensor] = []
    with torch.cuda.device(dev):
        group: list[tuple] = []  # (audio, n_samples, plan)
        n_win = 0

        def flush():
            nonlocal group, n_win
            if not group:
                return
            calls, pcm_off, frm_off = plan_packed_calls([g[1] for g in group], chunk_f, batch_size, step, fpw)
            pcm_all = torch.empty(int(pcm_off[-1]), dtype=torch.float32, device=dev)
            logits = torch.empty((int(frm_off[-1]), n_labels), dtype=torch.float32, device=dev)
            for k, (audio, n_s, _) in enumerate(group):
                if n_s > 0:
                    PcmSource(audio, dev, dev_out=pcm_all[int(pcm_off[k]): int(pcm_off[k + 1])]).all()
            for c in calls:
                tab = torch.tensor([c.sample_offsets, c.frame_offsets], dtype=torch.int64).to(dev, non_blocking=True)
                engine.forward_windows(pcm_all, tab[0], len(c.sample_offsets), c.win_len, logits, tab[1], c.frames_per_window)
            for k in range(len(group)):
                out.append(logits[int(frm_off[k]): int(frm_off[k + 1])])
            group, n_win = [], 0

        for a in audios:
            n_s = audio_n_samples(a)
            plan = plan_windows(n_s, chunk_f, batch_size, step, fpw)
            if group and n_win + plan.n_windows > PACK_MAX_WINDOWS:
                flush()
            group.append((a, n_s, plan))
            n_win += plan.n_windows
        flush()
    return out


def _lower_bounds(thresholds: dict, n_labels: int) -> list[float]:
    assert n_labels == len(thresholds)
    return [float(lab["lower_bound"]) for lab in thresholds.values()]


def apply_thresholds(feature_tensor: torch.Tensor, thresholds: dict[str, dict[str, float]], device="cuda") -> torch.Tensor:
    """``sigmoid(logits) > lower_bound`` per label (dict order = label order), as a bool tensor.
    Evaluated in the logit domain against cuts derived from torch's fp32 sigmoid, so the result is
    bit-identical to the reference's (inference.py:214-234)."""
    bounds = _lower_bounds(thresholds, feature_tensor.shape[-1])
    x = feature_tensor.to(_cuda_device(device), torch.float32).contiguous()
    with torch.cuda.device(x.device):
        return ops.threshold_mask(x, [logit_cut(t) for t in bounds], mode=ops.DECODE_LOGIT)


def _table_to_intervals(table: np.ndarray, conv_settings: ConvolutionSettings, labels) -> list[tuple[int, int, str]]:
    if table.shape[0] == 0:
        return []
    s_frame = table[:, 2].astype(np.int64) // FRAME_SAMPLES
    e_frame = table[:, 3].astype(np.int64) // FRAME_SAMPLES  # one past the last active frame
    starts = np.maximum(0, np.array([conv_settings.rf_start_i(int(s)) for s in s_frame], dtype=np.int64)) \
        if conv_settings != INFERENCE_SETTINGS else np.maximum(0, s_frame * FRAME_SAMPLES)
    ends = np.array([conv_settings.rf_end_i(int(e) - 1) + 1 for e in e_frame], dtype=np.int64) \
        if conv_settings != INFERENCE_SETTINGS else e_frame * FRAME_SAMPLES
    # plain Python ints / strs, as the reference returns them; tolist() first: iterating numpy scalars is 5x slower
    return list(zip(starts.tolist(), ends.tolist(), [labels[c] for c in table[:, 1].tolist()]))


def create_intervals(thresholded_features: torch.Tensor, conv_settings: ConvolutionSettings,
                     label_encoder: MultiLabelEncoder) -> list[tuple[int, int, str]]:
    """Per-label maximal runs of True -> ``(start_sample, end_sample, label)``, label-major
    (inference.py:237-263), extracted by the run-length kernel instead of NumPy + a Python loop."""
    m = torch.as_tensor(thresholded_features)
    if m.numel() == 0:
        return []
    x = m.to(m.device if m.is_cuda else _cuda_device("cuda"), torch.float32).contiguous()
    with torch.cuda.device(x.device):
        table = ops.decode_intervals(x, [0.5] * x.shape[-1], mode=ops.DECODE_LOGIT).cpu().numpy()
    return _table_to_intervals(table, conv_settings, label_encoder.base_labels)


def decode_logits(logits: torch.Tensor, thresholds: dict, label_encoder: MultiLabelEncoder,
                  conv_settings: ConvolutionSettings = INFERENCE_SETTINGS, file_offsets=None, *,
                  hysteresis: bool = False, max_gap_s: float = 0.0, min_duration_s: float = 0.0):
    """Fused ``apply_thresholds`` + ``create_intervals`` on device logits (one pass, no boolean tensor).
    With ``file_offsets`` the frames of several files are decoded at once; returns one list per file.

    Optional post-processing, all off by default (= the reference's behaviour): ``hysteresis`` uses each label's
    ``upper_bound`` as the onset and ``lower_bound`` as the offset threshold (an ``upper_bound`` of 1.0 or more,
    the reference's default, can never fire, so it is ignored); ``max_gap_s`` merges intervals of a label
    separated by at most that many seconds; ``min_duration_s`` drops shorter intervals."""
    bounds = _lower_bounds(thresholds, logits.shape[-1])
    if logits.shape[0] == 0:
        return [] if file_offsets is None else [[] for _ in range(len(file_offsets) - 1)]
    onset = None
    if hysteresis:
        ups = [float(lab.get("upper_bound", 1.0)) for lab in thresholds.values()]
        onset = [logit_cut(u) if u < 1.0 else logit_cut(lo) for u, lo in zip(ups, bounds)]
    with torch.cuda.device(logits.device):
        dev_table = ops.decode_intervals(logits.contiguous(), [logit_cut(t) for t in bounds], file_offsets=file_offsets,
                                         mode=ops.DECODE_LOGIT, onset=onset)
        if max_gap_s > 0.0 or min_duration_s > 0.0:
            dev_table = ops.postprocess_intervals(dev_table.contiguous(), int(round(max_gap_s * 16_000)),
                                                  int(round(min_duration_s * 16_000)))
        table = dev_table.cpu().numpy()
    labels = label_encoder.base_labels
    if file_offsets is None:
        return _table_to_intervals(table, conv_settings, labels)
    return [_table_to_intervals(table[table[:, 0] == f], conv_settings, labels) for f in range(len(file_offsets) - 1)]


def write_intervals(intervals: list[tuple[int, int, str]], audio_path: Path, output_p: Path) -> None:
    """``output_p/raw_rttm/<stem>.rttm`` in the reference's text format (inference.py:266-283)."""
    rttm_out = Path(output_p) / "raw_rttm"
    rttm_out.mkdir(exist_ok=True, parents=True)
    uri = Path(audio_path).stem
    with (rttm_out / f"{uri}.rttm").open("w") as f:
        f.write("".join(rttm_line(uri, s, e, lab) + "\n" for s, e, lab in intervals))


def default_thresholds(label_encoder: MultiLabelEncoder) -> dict:
    return {label: {"lower_bound": 0.5, "upper_bound": 1.0} for label in label_encoder._labels}


_D2H_STREAMS: dict[int, "torch.cuda.Stream"] = {}


def _d2h_stream(dev: torch.device) -> "torch.cuda.Stream":
    if dev.index not in _D2H_STREAMS:
        _D2H_STREAMS[dev.index] = torch.cuda.Stream(device=dev)
    return _D2H_STREAMS[dev.index]


class _PinnedPool:
    """Page-locked host buffers for the per-file read-backs, recycled by size class: ``cudaHostAlloc`` costs a fraction
    of a millisecond, which a corpus of short clips would otherwise pay three times per file."""

    def __init__(self):
        self._free: dict[int, list[torch.Tensor]] = {}

    def take(self, nbytes: int) -> torch.Tensor:
        size = 1 << max(int(nbytes - 1).bit_length(), 8)
        bucket = self._free.setdefault(size, [])
        return bucket.pop() if bucket else torch.empty(size, dtype=torch.uint8).pin_memory()

    def give(self, buf: torch.Tensor) -> None:
        self._free.setdefault(buf.numel(), []).append(buf)


_PINNED = _PinnedPool()


class _FileJob:
    """One file in flight: everything is queued on the device, nothing has been waited for yet.  The interval table
    (worst-case sized, a few MB per hour of audio), its row count and -- if asked for -- the logits travel to pinned
    host memory on a side stream, so the main stream never stalls on a per-file read-back."""

    def __init__(self, audio_path, model, config, batch_size, device, thresholds, save_logits, window_step, slot_base=0):
        self.stem = Path(audio_path).stem if not isinstance(audio_path, (np.ndarray, torch.Tensor)) else "audio"
        self.model = model
        dev = _cuda_device(device)
        logits = apply_model_on_audio(audio_path=audio_path, model=model, batch_size=batch_size,
                                      chunk_duration_s=config.audio.chunk_duration_s, conv_settings=INFERENCE_SETTINGS,
                                      device=dev, window_step=window_step, slot_base=slot_base)
        bounds = _lower_bounds(thresholds, logits.shape[-1])
        with torch.cuda.device(dev):
            self.table, self.count = ops.decode_intervals_async(logits, [logit_cut(t) for t in bounds], mode=ops.DECODE_LOGIT)
            main, side = torch.cuda.current_stream(dev), _d2h_stream(dev)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                self._bufs = [_PINNED.take(4), _PINNED.take(max(self.table.numel(), 1) * 4)]
                self.host_count = self._bufs[0][:4].view(torch.int32)
                self.host_table = self._bufs[1][: self.table.numel() * 4].view(torch.int32).view(self.table.shape)
                self.host_count.copy_(self.count, non_blocking=True)
                self.host_table.copy_(self.table, non_blocking=True)
                self.host_logits = None
                if save_logits:
                    self._bufs.append(_PINNED.take(max(logits.numel(), 1) * 4))
                    self.host_logits = self._bufs[2][: logits.numel() * 4].view(torch.float32).view(logits.shape)
                    self.host_logits.copy_(logits, non_blocking=True)
                self.done = torch.cuda.Event()
                self.done.record(side)
            for t in (logits, self.table, self.count):
                t.record_stream(side)

    def finish(self, output_p) -> list[tuple[int, int, str]]:
        """Wait for this file's copies only, then format: RTTM (+ logits file), as inference.py:333-357."""
        self.done.synchronize()
        le = self.model.label_encoder
        n = int(self.host_count[0])
        intervals = _table_to_intervals(self.host_table[:n].numpy(), INFERENCE_SETTINGS, le.base_labels)
        if self.host_logits is not None:
            logits_out_p = Path(output_p) / "logits"
            logits_out_p.mkdir(parents=True, exist_ok=True)
            torch.save({le.inv_transform(i): self.host_logits[:, i].clone() for i in range(le.n_labels)},
                       f"{logits_out_p}/{self.stem}-logits_dict_t.pt")
        if output_p is not None:
            write_intervals(intervals=intervals, audio_path=Path(self.stem), output_p=output_p)
        for buf in self._bufs:  # everything has been copied out of the pinned buffers (numpy -> Python ints, clone)
            _PINNED.give(buf)
        self._bufs, self.host_table, self.host_count, self.host_logits = [], None, None, None
        return intervals


def infer_file(audio_path, model: BaseSegmentationModel, output_p: Path, config: Config, batch_size: int,
               device="cuda", thresholds: None | dict = None, save_logits: bool = False, window_step: int | None = None):
    """Window, forward, threshold, decode and write one file (inference.py:286-357). Returns the intervals."""
    if thresholds is None:
        thresholds = default_thresholds(model.label_encoder)
    job = _FileJob(audio_path, model, config, batch_size, device, thresholds, save_logits, window_step)
    return job.finish(output_p)


#: units a rank keeps queued on its GPU before it claims the next one (dynamic walk of `infer_corpus`)
DYNAMIC_IN_FLIGHT = int(os.environ.get("SEGMA_DYNAMIC_IN_FLIGHT", "3"))
_CORPUS_WALKS = 0


def infer_corpus(audios, model: BaseSegmentationModel, config: Config, batch_size: int = 128, device="cuda",
                 thresholds: None | dict = None, shard: tuple[int, int] | None = None, sizes=None,
                 window_step: int | None = None, gather: bool = True) -> torch.Tensor:
    """A whole corpus -> one int32 ``(n_intervals, 4)`` device table ``(file, label, start_sample, end_sample)`` with
    ``file`` the index into ``audios``, ordered by file, then label, then time.

    The multi-GPU form of the reference's serial file loop (inference.py:442-458): with
    ``shard=(rank, world_size)`` the work is partitioned by audio file and by window batch -- files are assigned longest
    first, and a file that alone would unbalance the ranks is cut into contiguous ranges of its forward calls
    (``geometry.plan_work_units``; batch boundaries never move, they are part of the result).  Nothing is exchanged while
    the units run and nothing is read back per unit: every table stays on the device at its worst-case size with its row
    count.  At the very end the counts are read once, the tables compacted, shifted to their place on the file timeline
    and all-gathered once (NCCL); runs that cross a cut are fused by the gap-0 interval merge
    (``distributed.merge_split_files``).  ``audios`` holds paths or 1-D arrays."""
    from .distributed import gather_corpus_tables
    from .io import audio_n_samples

    if thresholds is None:
        thresholds = default_thresholds(model.label_encoder)
    dev = _cuda_device(device)
    chunk_f = int(config.audio.chunk_duration_s * 16_000)
    fpw = model.n_keep if model.family == "whisper" else conv_frames(chunk_f)
    world = shard[1] if shard is not None else 1
    if sizes is None:
        sizes = [audio_n_samples(a) for a in audios]
    step = Chunkyfier(batch_size, chunk_f, INFERENCE_SETTINGS).step if window_step is None else int(window_step)
    units = plan_work_units(list(sizes), world if window_step is None else 1, chunk_f, batch_size, step, fpw)
    pack = model.family == "wav2vec2" and window_step is None and os.environ.get("SEGMA_PACK_FILES", "1") != "0"
    # SEGMA_DYNAMIC_SHARD=1: units are handed out on demand (distributed.UnitQueue) instead of in static shares: a rank
    # keeps at most DYNAMIC_IN_FLIGHT units queued on its GPU and claims the next one, longest first, only when one has
    # finished, so that faster boards take more units.  Off by default: on the 256-file corpus over 8 B200 the static
    # longest-first shares were 1 % faster (19.9 against 19.6-19.8 audio-h/s, profiles/r02z_dynamic_shard_ab_8gpu.txt) --
    # waiting on a unit's completion event before the next claim costs more host-side pipelining than the few percent
    # of board-to-board spread give back at that size.
    dynamic = (shard is not None and world > 1 and not pack and os.environ.get("SEGMA_DYNAMIC_SHARD", "0") == "1"
               and torch.distributed.is_available() and torch.distributed.is_initialized())
    if dynamic:
        from .distributed import UnitQueue

        global _CORPUS_WALKS
        _CORPUS_WALKS += 1
        order = sorted(units, key=lambda u: (-u.n_windows, u.file, u.batch_lo))
        queue = UnitQueue(f"corpus{_CORPUS_WALKS}", len(order))
        mine = []
    else:
        mine = assign_units(units, world)[shard[0]] if shard is not None else units
    any_split = any(not u.whole_file for u in units)
    cuts = [logit_cut(t) for t in _lower_bounds(thresholds, model.label_encoder.n_labels)]
    results: dict[int, tuple] = {}  # position in `mine` -> (table, count)
    if pack:
        # independent windows: whole files are packed across file boundaries into full forward calls
        whole = [k for k, u in enumerate(mine) if u.whole_file]
        packed = apply_model_on_audios([audios[mine[k].file] for k in whole], model, INFERENCE_SETTINGS, dev,
                                       batch_size=batch_size, chunk_duration_s=config.audio.chunk_duration_s)
        for k, logits in zip(whole, packed):
            with torch.cuda.device(dev):
                results[k] = ops.decode_intervals_async(logits.contiguous(), cuts, mode=ops.DECODE_LOGIT)
    # Everything else runs unit by unit.  Units are independent of each other: short ones, whose forward calls are too
    # small to fill 148 SMs, take turns on FILE_STREAMS streams, each with its own workspace slot; long ones run alone
    # on the main stream.
    main = torch.cuda.current_stream(dev)
    lanes = _side_streams(dev, FILE_STREAMS) if FILE_STREAMS > 1 else []
    for st in lanes:
        st.wait_stream(main)
    turn = 0
    in_flight: list = []  # dynamic walk: completion events of the units queued on this GPU

    def walk():
        if not dynamic:
            yield from enumerate(mine)
            return
        while True:
            while len(in_flight) >= DYNAMIC_IN_FLIGHT:
                in_flight.pop(0).synchronize()
            i = queue.claim()
            if i is None:
                return
            mine.append(order[i])
            yield len(mine) - 1, order[i]

    for k, u in walk():
        if k in results:
            continue
        small = bool(lanes) and u.n_windows < SMALL_FILE_WINDOWS
        lane = turn % FILE_STREAMS if small else -1
        turn += 1 if small else 0
        with torch.cuda.stream(lanes[lane] if small else main):
            logits = apply_model_on_audio(audios[u.file], model, INFERENCE_SETTINGS, dev, batch_size=batch_size,
                                          chunk_duration_s=config.audio.chunk_duration_s, window_step=window_step,
                                          slot_base=(1 + lane) * max(N_STREAMS, 1) if small else 0,
                                          batch_range=None if u.whole_file else (u.batch_lo, u.batch_hi))
            with torch.cuda.device(dev):
                table, count = ops.decode_intervals_async(logits, cuts, mode=ops.DECODE_LOGIT)
        if small:  # allocated on the lane's stream, consumed by the final exchange on the main stream
            table.record_stream(main)
            count.record_stream(main)
        results[k] = (table, count)
        if dynamic:
            done = torch.cuda.Event()
            done.record(lanes[lane] if small else main)
            in_flight.append(done)
    for st in lanes:
        main.wait_stream(st)
    if dynamic:  # claimed in longest-first order: back to (file, batch) order like the static shares
        perm = sorted(range(len(mine)), key=lambda k: (mine[k].file, mine[k].batch_lo))
        results = {j: results[k] for j, k in enumerate(perm)}
        mine = [mine[k] for k in perm]
    offsets = []
    for u in mine:  # first sample of the unit on its file's timeline
        if u.whole_file:
            offsets.append(0)
        else:
            plan = plan_windows(sizes[u.file], chunk_f, batch_size, step, fpw)
            offsets.append(batch_frame_range(plan, u.batch_lo, u.batch_hi)[0] * FRAME_SAMPLES)
    with torch.cuda.device(dev):
        return gather_corpus_tables([u.file for u in mine], [results[k][0] for k in range(len(mine))],
                                    [results[k][1] for k in range(len(mine))], device=dev,
                                    gather=gather and shard is not None and shard[1] > 1, sample_offsets=offsets,
                                    merge_split=any_split and gather)


def get_list_of_files_to_process(wavs: Path, recursive: bool = False, uris: Path | None = None) -> tuple[list[Path], int]:
    """Sorted list of ``.wav`` files under ``wavs`` (or those named in ``uris``) (inference.py:360-395)."""
    wavs = Path(wavs)
    if not wavs.exists():
        raise FileNotFoundError(f"Path `{wavs=}` does not exists")
    if uris:
        with Path(uris).open("r") as f:
            files = [(wavs / line.strip()).with_suffix(".wav") for line in f.readlines()]
    elif recursive:
        import warnings

        warnings.warn("Search for .wav files is done recursively, might be slow.")
        files = list(wavs.rglob("*.wav"))
    else:
        files = list(wavs.glob("*.wav"))
    return sorted(files), len(files)


def run_inference_on_audios(config, uris, wavs, checkpoint, output, thresholds, batch_size: int,
                            device: Literal["gpu", "cuda"] = "cuda", recursive: bool = False, save_logits: bool = False,
                            logger: Logger | None = None, shard: tuple[int, int] | None = None) -> list[Path]:
    """File list -> per-file RTTM (inference.py:398-459).  ``shard=(rank, world_size)`` restricts the loop
    to this rank's files (see ``segma_b200.distributed``); the default processes every file."""
    wavs, checkpoint, output = Path(wavs), Path(checkpoint), Path(output)
    device = "cuda" if device == "gpu" else device
    if not checkpoint.exists():
        raise ValueError(f"Path `{checkpoint=}` does not exists")
    if thresholds:
        if not Path(thresholds).exists():
            raise ValueError("Path to a valid threshold dict does not exist.")
        with Path(thresholds).open("r") as f:
            thresholds = yaml.safe_load(f)
    files, n_files = get_list_of_files_to_process(wavs, recursive, uris)
    cfg: Config = load_config(config) if not isinstance(config, Config) else config
    if "hydra" not in cfg.model.name:
        raise ValueError("only MultiLabelEncoder is supported")
    l_encoder = MultiLabelEncoder(labels=cfg.data.classes)
    model = Models[cfg.model.name].load_from_checkpoint(checkpoint_path=checkpoint, label_encoder=l_encoder, config=cfg,
                                                        train=False)
    model.eval()
    model.to(torch.device(device))
    mine = files
    if shard is not None:
        from .distributed import assign_files

        sizes = [get_audio_info(p).n_samples for p in files]
        mine = [files[i] for i in assign_files(sizes, shard[1])[shard[0]]]
    if not thresholds:
        thresholds = default_thresholds(model.label_encoder)
    # Files are queued back to back: file i's interval table comes back on a side stream and is formatted on the host
    # while the device already runs file i + 1 ... i + MAX_FILES_IN_FLIGHT (the reference loop is serial, 442-458).
    # Short files (forward calls too small to fill the GPU) take turns on FILE_STREAMS streams, as in infer_corpus.
    from .io import audio_n_samples

    dev = _cuda_device(device)
    main = torch.cuda.current_stream(dev)
    lanes = _side_streams(dev, FILE_STREAMS) if FILE_STREAMS > 1 else []
    for st in lanes:
        st.wait_stream(main)
    small_below = SMALL_FILE_WINDOWS * cfg.audio.chunk_duration_f
    in_flight: list[_FileJob] = []
    turn = 0
    for i, audio_path in enumerate(mine, 1):
        s = f"({i:>{len(str(n_files))}}/{n_files}) - running inference for file: '{audio_path.stem}'"
        if logger:
            logger.info(s)
        else:
            print(f"[log] - {s}", flush=True)
        small = bool(lanes) and audio_n_samples(audio_path) < small_below
        lane = turn % FILE_STREAMS if small else -1
        turn += 1 if small else 0
        with torch.cuda.stream(lanes[lane] if small else main):
            in_flight.append(_FileJob(audio_path, model, cfg, batch_size, device, thresholds, save_logits, None,
                                      slot_base=(1 + lane) * max(N_STREAMS, 1) if small else 0))
        while len(in_flight) > max(MAX_FILES_IN_FLIGHT, FILE_STREAMS):
            in_flight.pop(0).finish(output)
    for job in in_flight:
        job.finish(output)
    for st in lanes:
        main.wait_stream(st)
    return mine


def main(argv=None):
    parser = argparse.ArgumentParser(description="segma_b200 sliding-window inference (same flags as segma.inference)")
    parser.add_argument("--config", type=str, required=True)
    parser.add_argument("--uris")
    parser.add_argument("--wavs", required=True, default="data/debug/wav")
    parser.add_argument("--checkpoint", default="models/last/best.ckpt")
    parser.add_argument("--output", required=True)
    parser.add_argument("--thresholds")
    parser.add_argument("--batch_size", default=128, type=int)
    parser.add_argument("--device", default="cuda", choices=["gpu", "cuda"])
    args = parser.parse_args(argv)
    run_inference_on_audios(**vars(args))


if __name__ == "__main__":
    main()
