"""Frame/sample geometry: the reference's own ConvolutionSettings known-answer tests
(/root/reference/tests/test_ConvolutionSettings.py:4-28), the golden values produced by the reference's
class, and the window/batch plan against the oracle's restatement of inference.py:129-206."""
from pathlib import Path

import numpy as np
import pytest

from oracle import segma_oracle as O
from segma_b200.geometry import INFERENCE_SETTINGS, Chunkyfier, ConvolutionSettings, conv_frames, plan_windows

GOLDEN = Path(__file__).parent / "golden"


def test_rf_start_i():
    assert ConvolutionSettings((3, 2), (3, 1), (1, 0)).rf_start_i(0) == -1
    assert ConvolutionSettings((2,), (1,), (0,)).rf_start_i(0) == 0


def test_rf_end_i():
    assert ConvolutionSettings((3, 2), (3, 1), (1, 0)).rf_end_i(0) == 4
    assert ConvolutionSettings((2,), (1,), (0,)).rf_end_i(0) == 1


def test_rf_size():
    assert ConvolutionSettings((3, 2), (3, 1), (1, 0)).rf_size == 6
    assert ConvolutionSettings((2,), (1,), (0,)).rf_size == 2


def test_mismatched_settings_raise():
    with pytest.raises(ValueError):
        ConvolutionSettings((3, 2), (3,), (1, 0))


def test_golden_receptive_fields():
    g = np.load(GOLDEN / "geometry.npz")
    for row in g["rf"]:
        L = int(row[0])
        ks, ss, ps = tuple(row[1:1 + L]), tuple(row[1 + L:1 + 2 * L]), tuple(row[1 + 2 * L:1 + 3 * L])
        u, start, end, size, step = (int(v) for v in row[1 + 3 * L:6 + 3 * L])
        cs = ConvolutionSettings(tuple(int(k) for k in ks), tuple(int(s) for s in ss), tuple(int(p) for p in ps))
        assert (cs.rf_start_i(u), cs.rf_end_i(u), cs.rf_size, cs.rf_step) == (start, end, size, step)
        assert O.rf_start(u, ks, ss, ps) == start and O.rf_end(u, ks, ss, ps) == end and O.rf_size(ks, ss) == size
    whisper = ConvolutionSettings((400, 3, 3), (160, 1, 2), (200, 1, 1))
    for chunk, strict_n, loose_n in g["n_windows"]:
        assert INFERENCE_SETTINGS.n_windows(int(chunk), True) == strict_n
        assert whisper.n_windows(int(chunk), False) == loose_n


def test_chunkyfier_matches_reference_constants():
    c = Chunkyfier(128, 64000, INFERENCE_SETTINGS)
    assert (c.n_windows, c.missing_n_frames, c.step) == (199, 320, 63680)
    assert c.chunk_start_i(3) == 3 * 63680 and c.chunk_end_i(3) == 3 * 63680 + 64000
    assert c.batch_start_i(2) == 2 * 128 * 63680
    assert c.batch_end_i_coverage(0) == 128 * 63680
    assert c.get_n_fitting_chunks(64000) == 1 and c.get_n_fitting_chunks(63999) == 0
    assert c.get_n_fitting_chunks(57_600_000) == 904


@pytest.mark.parametrize("n", [0, 399, 400, 50_000, 64_000, 64_319, 127_460, 127_680, 193_234, 960_000, 2_880_000, 57_600_000])
@pytest.mark.parametrize("bs", [1, 2, 128])
def test_plan_equals_oracle_batches(n, bs):
    plan = plan_windows(n, 64000, bs)
    mine = [(b.start_sample, b.n_windows, b.win_len) for b in plan.batches]
    assert mine == O.file_batches(n, 64000, bs)
    assert plan.n_frames == conv_frames(n)  # contiguous 20 ms grid: (n-400)//320+1
    assert sum(b.n_windows * b.frames_per_window for b in plan.batches) == plan.n_frames


def test_one_hour_file_counts():
    plan = plan_windows(57_600_000)
    assert plan.n_windows == 905 and plan.n_frames == 179_999
    assert [b.n_windows for b in plan.batches] == [128] * 7 + [8, 1]
    assert plan.batches[-1].is_tail and plan.batches[-1].win_len == 33_280 and plan.batches[-1].frames_per_window == 103


@pytest.mark.parametrize("win_s,overlap", [(2, 0.5), (3, 0.75), (4, 0.9), (6, 0.5), (8, 0.75)])
def test_overlapping_plans_cover_every_frame(win_s, overlap):
    win = win_s * 16000
    F = conv_frames(win)
    step = max(320, int(win * (1 - overlap)) // 320 * 320)
    n = 1_000_000
    plan = plan_windows(n, win, 16, step, F)
    cover = np.zeros(plan.n_frames, dtype=int)
    for b in plan.batches:
        for i in range(b.n_windows):
            off = (b.first_window + i) * plan.step_frames
            cover[off: off + b.frames_per_window] += 1
    assert cover.min() >= 1
    assert plan.n_frames <= conv_frames(n)


def test_bad_steps_rejected():
    with pytest.raises(ValueError):
        plan_windows(100_000, 64000, 4, step=1000)
    with pytest.raises(ValueError):
        plan_windows(100_000, 64000, 4, step=64000 + 320)


def test_packed_calls_keep_every_files_windows():
    """Cross-file packing (models whose windows are independent): each file keeps exactly the windows and frames of
    ``plan_windows``; full windows of all files fill calls of ``batch_size``, tails of equal length share calls."""
    from segma_b200.geometry import plan_packed_calls, plan_windows

    lens = [64000 + 63680 * 2 + 9000, 300, 64000, 70_000, 64000 + 63680 + 9000, 5000, 0, 70_000, 12_345]
    calls, pcm_off, frm_off = plan_packed_calls(lens, 64_000, 3)
    assert pcm_off == [0] + list(np.cumsum(lens)) and len(frm_off) == len(lens) + 1
    # every frame of every file is written exactly once
    hits = np.zeros(frm_off[-1], dtype=int)
    for c in calls:
        assert 0 < len(c.sample_offsets) <= 3 and len(c.sample_offsets) == len(c.frame_offsets)
        for w, f in zip(c.sample_offsets, c.frame_offsets):
            hits[f: f + c.frames_per_window] += 1
            k = int(np.searchsorted(pcm_off, w, side="right")) - 1
            assert w + c.win_len <= pcm_off[k + 1]  # a window never crosses into the next file
            assert frm_off[k] <= f and f + c.frames_per_window <= frm_off[k + 1]
    assert (hits == 1).all()
    for k, n in enumerate(lens):
        assert frm_off[k + 1] - frm_off[k] == plan_windows(n, 64_000, 3).n_frames == (max((n - 400) // 320 + 1, 0) if n >= 400 else 0)
    full = [c for c in calls if c.win_len == 64_000]
    assert sum(len(c.sample_offsets) for c in full) == sum(max((n - 64000) // 63680 + 1, 0) if n >= 64000 else 0 for n in lens)
    assert all(len(c.sample_offsets) == 3 for c in full[:-1])  # packed across file boundaries
    tails = [c for c in calls if c.win_len != 64_000]
    assert any(len(c.sample_offsets) == 2 for c in tails)  # the two 9000-sample tails (and the two 6320-sample ones) share a call


def test_work_units_cover_every_batch_once_and_balance():
    """Multi-GPU partition by audio file and window batch: the units of a file are contiguous batch ranges that cover
    its plan exactly once; small files stay whole; one very long file no longer pins the slowest rank."""
    from segma_b200.geometry import assign_units, batch_frame_range, plan_work_units

    hour = 57_600_000
    sizes = [hour * 6, hour // 60, 300, 64000, hour // 10, 0, 63680 * 128 + 320]
    for world in (1, 2, 4, 8):
        units = plan_work_units(sizes, world, 64000, 128, 63680, 199)
        for f, n in enumerate(sizes):
            plan = plan_windows(n, 64000, 128, 63680, 199)
            mine = [u for u in units if u.file == f]
            assert [u.batch_lo for u in mine] == sorted(u.batch_lo for u in mine)
            assert mine[0].batch_lo == 0 and mine[-1].batch_hi == len(plan.batches)
            assert all(a.batch_hi == b.batch_lo for a, b in zip(mine, mine[1:]))
            assert sum(u.n_windows for u in mine) == plan.n_windows
            assert all(u.whole_file == (len(mine) == 1) for u in mine)
            if len(mine) > 1:  # frame ranges of the pieces tile the file's frame grid
                ranges = [batch_frame_range(plan, u.batch_lo, u.batch_hi) for u in mine]
                assert ranges[0][0] == 0 and ranges[-1][1] == plan.n_frames
                assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        parts = assign_units(units, world)
        assert sorted((u.file, u.batch_lo) for p in parts for u in p) == sorted((u.file, u.batch_lo) for u in units)
        assert assign_units(units, world) == parts  # deterministic
        loads = [sum(u.n_windows for u in p) for p in parts]
        if world > 1:
            assert len([u for u in units if u.file == 0]) > world  # the 6 h file is cut
            assert max(loads) <= 1.15 * (sum(loads) / world), loads  # whole files would give world x the mean
    # a corpus of equal files is left alone: whole files, 125 per rank
    units = plan_work_units([hour] * 1000, 8, 64000, 128, 63680, 199)
    assert all(u.whole_file for u in units) and [len(p) for p in assign_units(units, 8)] == [125] * 8
    # one file, eight ranks: every rank gets work
    units = plan_work_units([hour], 8, 64000, 128, 63680, 199)
    assert all(len(p) >= 1 for p in assign_units(units, 8))
