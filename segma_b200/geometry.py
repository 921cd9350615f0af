"""Frame <-> sample geometry of the sliding-window path (integer arithmetic only).

Mirrors the reference's ``ConvolutionSettings`` (/root/reference/src/segma/models/base.py:19-142)
and ``Chunkyfier`` (/root/reference/src/segma/inference.py:21-89) so that the drop-in
``apply_model_on_audio`` cuts files into exactly the reference's windows and batches,
and generalises them to overlapping windows (``WindowPlan``).
"""
from __future__ import annotations

from dataclasses import dataclass
from functools import cached_property
from itertools import accumulate
from operator import mul


@dataclass
class ConvolutionSettings:
    """Receptive-field arithmetic of a stack of 1-D convolutions.

    Same public surface as the reference class (models/base.py:19-142): ``rf_start_i``,
    ``rf_end_i``, ``rf_size``, ``rf_center_i``, ``rf_step``, ``n_windows``.
    """

    kernels: tuple[int, ...]
    strides: tuple[int, ...]
    paddings: tuple[int, ...]

    def __post_init__(self):
        if not (len(self.kernels) == len(self.strides) == len(self.paddings)):
            raise ValueError(
                "Given settings do not match, please provide matching dimensions for kernels, strides and paddings."
            )

    # jump of layer l = product of the strides below it (1 for the first layer)
    def _jumps(self) -> list[int]:
        return list(accumulate((1,) + tuple(self.strides[:-1]), mul))

    def _total_stride(self) -> int:
        s = 1
        for v in self.strides:
            s *= v
        return s

    def _pad_offset(self) -> int:
        return sum(p * j for p, j in zip(self.paddings, self._jumps()))

    def rf_start_i(self, u_L: int) -> int:
        """First input sample seen by output index ``u_L`` (may be negative). base.py:31-50."""
        return u_L * self._total_stride() - self._pad_offset()

    def rf_end_i(self, v_L: int) -> int:
        """Last input sample (inclusive) seen by output index ``v_L``. base.py:52-74."""
        back = sum((1 + p - k) * j for k, p, j in zip(self.kernels, self.paddings, self._jumps()))
        return v_L * self._total_stride() - back

    @cached_property
    def rf_size(self) -> int:
        """Receptive-field size in input samples. base.py:76-91."""
        return 1 + sum((k - 1) * j for k, j in zip(self.kernels, self._jumps()))

    def rf_center_i(self, u_L: int):
        """Centre of the receptive field (float). base.py:93-103."""
        return u_L * self._total_stride() + (self.rf_size - 1) / 2 - self._pad_offset()

    @cached_property
    def rf_step(self) -> int:
        """Distance in samples between two consecutive receptive fields. base.py:105-117."""
        d_start = abs(self.rf_start_i(0) - self.rf_start_i(1))
        d_end = abs(self.rf_end_i(0) - self.rf_end_i(1))
        d_mid = abs(self.rf_center_i(0) - self.rf_center_i(1))
        assert d_start == d_end == d_mid
        return d_start

    def n_windows(self, chunk_duration_f: int, strict: bool = True) -> int:
        """Number of output frames for a chunk of ``chunk_duration_f`` samples.

        Keeps the reference's "+1 to the step if any kernel is even" rule (base.py:131-142),
        including its off-by-one for >= 7 s windows.
        """
        step = int(self.rf_step + (1 if any(k % 2 == 0 for k in self.kernels) else 0))
        if strict:
            return (chunk_duration_f - self.rf_size) // step + 1
        return chunk_duration_f // step


#: the settings ``infer_file`` decodes intervals with (inference.py:315-319)
INFERENCE_SETTINGS = ConvolutionSettings(kernels=(320,), strides=(320,), paddings=(0,))

FRAME_SAMPLES = 320  # 20 ms output frame
MIN_TAIL_SAMPLES = 400  # shortest tail the reference still forwards (inference.py:195)


def conv_frames(n_samples: int) -> int:
    """Frames a waveform model emits for ``n_samples``: ``(n-400)//320 + 1``."""
    return 0 if n_samples < MIN_TAIL_SAMPLES else (n_samples - MIN_TAIL_SAMPLES) // FRAME_SAMPLES + 1


class Chunkyfier:
    """Window / batch index arithmetic with the reference's method names (inference.py:21-89).

    The reference hard-asserts 199 frames and 320 missing samples (4 s windows only);
    here the same two quantities are derived for any window length and the assertion
    is relaxed to "the missing part is a whole number of frames".
    """

    def __init__(self, batch_size: int, chunk_duration_f: int, cnn_settings: ConvolutionSettings):
        self.cnn_settings = cnn_settings
        self.chunk_duration_f = chunk_duration_f
        self.batch_size = batch_size
        self.n_windows = cnn_settings.n_windows(chunk_duration_f, strict=True)
        self.missing_n_frames = chunk_duration_f - self.n_windows * cnn_settings.rf_step
        assert self.n_windows > 0 and self.missing_n_frames > 0

    @property
    def step(self) -> int:
        return self.chunk_duration_f - self.missing_n_frames

    def chunk_start_i(self, i: int) -> int:
        return i * self.step

    def chunk_end_i(self, i: int) -> int:
        return self.chunk_start_i(i) + self.chunk_duration_f

    def chunk_end_i_coverage(self, i: int) -> int:
        return (i + 1) * self.step

    def batch_start_i(self, i: int) -> int:
        return i * self.batch_size * self.step

    def batch_end_i(self, i: int) -> int:
        return self.batch_start_i(i) + self.batch_size * self.chunk_duration_f

    def batch_end_i_coverage(self, i: int) -> int:
        return self.batch_end_i(i) - self.batch_size * self.missing_n_frames

    def get_n_fitting_chunks(self, n_frames: int) -> int:
        """Complete windows (with the reference's 320-sample overlap) that fit in ``n_frames`` samples."""
        return (n_frames - self.chunk_duration_f) // self.step + 1


@dataclass(frozen=True)
class WindowBatch:
    """One forward call of the reference loop: ``n_windows`` consecutive windows of one file."""

    first_window: int  # index of the first window in the file
    n_windows: int
    start_sample: int  # first PCM sample the batch reads
    win_len: int  # samples per window in this batch (tail: shorter)
    frames_per_window: int  # frames kept per window
    is_tail: bool = False


@dataclass(frozen=True)
class WindowPlan:
    """Every window and batch ``apply_model_on_audio`` forwards for a file of ``n_samples``.

    ``step`` defaults to the reference's ``win_len - 320`` (windows then tile the 20 ms frame
    grid and stitching is concatenation, inference.py:148-152,209-211); smaller steps
    (multiples of 320) give overlapping windows that are averaged in the logit domain.
    Batch boundaries follow inference.py:138-206: ``batch_size`` consecutive windows,
    then one remainder batch, then the variable-length tail alone.
    """

    n_samples: int
    win_len: int
    step: int
    frames_per_window: int
    batch_size: int
    batches: tuple[WindowBatch, ...]
    n_frames: int  # frames on the file timeline

    @property
    def n_windows(self) -> int:
        return sum(b.n_windows for b in self.batches)

    @property
    def step_frames(self) -> int:
        return self.step // FRAME_SAMPLES


def plan_windows(
    n_samples: int,
    win_len: int = 64_000,
    batch_size: int = 128,
    step: int | None = None,
    frames_per_window: int | None = None,
) -> WindowPlan:
    if step is None:
        step = win_len - FRAME_SAMPLES
    if step <= 0 or step % FRAME_SAMPLES != 0:
        raise ValueError(f"window step must be a positive multiple of {FRAME_SAMPLES} samples, got {step}")
    if batch_size <= 0:
        raise ValueError("batch_size must be positive")
    if frames_per_window is None:
        frames_per_window = conv_frames(win_len)
    if step > frames_per_window * FRAME_SAMPLES:
        raise ValueError("window step leaves uncovered frames between windows")

    batches: list[WindowBatch] = []
    n_fit = 0 if n_samples < win_len else (n_samples - win_len) // step + 1
    w = 0
    while n_fit - w >= batch_size:
        batches.append(WindowBatch(w, batch_size, w * step, win_len, frames_per_window))
        w += batch_size
    if n_fit - w > 0:
        batches.append(WindowBatch(w, n_fit - w, w * step, win_len, frames_per_window))
        w = n_fit
    n_frames = 0 if n_fit == 0 else (n_fit - 1) * (step // FRAME_SAMPLES) + frames_per_window
    tail_start = n_fit * step
    tail_len = n_samples - tail_start
    if tail_len >= MIN_TAIL_SAMPLES:
        f_tail = min(conv_frames(tail_len), frames_per_window)
        batches.append(WindowBatch(n_fit, 1, tail_start, tail_len, f_tail, is_tail=True))
        n_frames = max(n_frames, n_fit * (step // FRAME_SAMPLES) + f_tail)
    return WindowPlan(n_samples, win_len, step, frames_per_window, batch_size, tuple(batches), n_frames)


@dataclass(frozen=True)
class PackedCall:
    """One forward call over windows of several files (models without coupling between the windows of a call)."""

    win_len: int  # samples per window (a tail length for tail calls)
    frames_per_window: int  # frames kept per window
    sample_offsets: tuple[int, ...]  # first sample of each window in the packed PCM buffer
    frame_offsets: tuple[int, ...]  # first frame of each window in the packed logits buffer


def plan_packed_calls(n_samples: list[int], win_len: int = 64_000, batch_size: int = 128, step: int | None = None,
                      frames_per_window: int | None = None):
    """Files laid end to end in one PCM buffer and one logits buffer -> ``(calls, pcm_offsets, frame_offsets)``.

    Every file keeps exactly the windows of ``plan_windows`` (the reference's, inference.py:129-206); only their
    grouping into forward calls changes: full windows of all files fill calls of ``batch_size`` in file order, tails of
    equal length share calls.  ``pcm_offsets[k]`` / ``frame_offsets[k]`` are file k's first sample / frame
    (both lists end with the totals)."""
    plans = [plan_windows(n, win_len, batch_size, step, frames_per_window) for n in n_samples]
    pcm_off, frm_off = [0], [0]
    for n, pl in zip(n_samples, plans):
        pcm_off.append(pcm_off[-1] + n)
        frm_off.append(frm_off[-1] + pl.n_frames)
    full_w, full_f, tails = [], [], {}
    for k, pl in enumerate(plans):
        for b in pl.batches:
            if b.is_tail:
                tails.setdefault((b.win_len, b.frames_per_window), []).append(
                    (pcm_off[k] + b.start_sample, frm_off[k] + b.first_window * pl.step_frames))
            else:
                for i in range(b.n_windows):
                    full_w.append(pcm_off[k] + b.start_sample + i * pl.step)
                    full_f.append(frm_off[k] + (b.first_window + i) * pl.step_frames)
    calls = []
    if full_w:
        fpw = plans[0].frames_per_window
        for i in range(0, len(full_w), batch_size):
            calls.append(PackedCall(win_len, fpw, tuple(full_w[i: i + batch_size]), tuple(full_f[i: i + batch_size])))
    for (wl, keep), items in tails.items():
        for i in range(0, len(items), batch_size):
            part = items[i: i + batch_size]
            calls.append(PackedCall(wl, keep, tuple(p[0] for p in part), tuple(p[1] for p in part)))
    return calls, pcm_off, frm_off


@dataclass(frozen=True)
class WorkUnit:
    """A contiguous range of forward calls (window batches) of one file: the atomic work item of the multi-GPU
    partition (batches are independent of one another -- the LSTM couples the windows *inside* a call only)."""

    file: int
    batch_lo: int
    batch_hi: int  # exclusive
    n_windows: int
    whole_file: bool


def batch_frame_range(plan: WindowPlan, lo: int, hi: int) -> tuple[int, int]:
    """Frames ``[f_lo, f_hi)`` of the file timeline that batches ``[lo, hi)`` of a tiled plan write."""
    assert plan.step_frames == plan.frames_per_window, "batch ranges need windows that tile the frame grid"
    bs = plan.batches[lo:hi]
    if not bs:
        return 0, 0
    f_lo = bs[0].first_window * plan.step_frames
    last = bs[-1]
    f_hi = plan.n_frames if last.is_tail else (last.first_window + last.n_windows) * plan.step_frames
    return f_lo, f_hi


def plan_work_units(n_samples: list[int], world_size: int, win_len: int = 64_000, batch_size: int = 128,
                    step: int | None = None, frames_per_window: int | None = None) -> list[WorkUnit]:
    """Files -> work units for ``world_size`` ranks.  A file stays whole unless it alone would unbalance the ranks
    (more than half of a rank's fair share of windows): then it is cut into contiguous batch ranges of about a quarter of
    that share.  Batch boundaries are never moved -- they are part of the result for the Whisper family."""
    plans = [plan_windows(n, win_len, batch_size, step, frames_per_window) for n in n_samples]
    total = sum(p.n_windows for p in plans)
    share = max(total / max(world_size, 1), 1.0)
    units: list[WorkUnit] = []
    for f, pl in enumerate(plans):
        nb = len(pl.batches)
        if world_size <= 1 or nb <= 1 or pl.n_windows <= share / 2:
            units.append(WorkUnit(f, 0, nb, pl.n_windows, True))
            continue
        per = max(1, round(share / 4 / batch_size))  # batches per unit
        for lo in range(0, nb, per):
            hi = min(lo + per, nb)
            units.append(WorkUnit(f, lo, hi, sum(b.n_windows for b in pl.batches[lo:hi]), False))
    return units


def assign_units(units: list[WorkUnit], world_size: int) -> list[list[WorkUnit]]:
    """Longest-processing-time-first assignment of work units to ranks (deterministic); each rank's units come back in
    (file, batch) order."""
    import heapq

    heap = [(0, r) for r in range(world_size)]
    heapq.heapify(heap)
    out: list[list[WorkUnit]] = [[] for _ in range(world_size)]
    for u in sorted(units, key=lambda u: (-u.n_windows, u.file, u.batch_lo)):
        load, r = heapq.heappop(heap)
        out[r].append(u)
        heapq.heappush(heap, (load + max(u.n_windows, 1), r))
    return [sorted(v, key=lambda u: (u.file, u.batch_lo)) for v in out]
