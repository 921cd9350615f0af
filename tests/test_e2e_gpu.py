"""End-to-end parity of the Whisper-family path on the GPU against the oracle (the reference's fp32
arithmetic restated in oracle/segma_oracle.py and pinned against the reference itself).

Stated tolerances (north_star): interval decoding bit-exact on identical logits; logits within (a fraction of) a bf16
tolerance of the fp32 oracle with >= 99.9 % frame-label agreement."""
import numpy as np
import pytest
import torch

from oracle import segma_oracle as O
from segma_b200 import synth
from segma_b200.config import make_config
from segma_b200.encoders import MultiLabelEncoder
from segma_b200.inference import apply_model_on_audio, apply_thresholds, create_intervals, decode_logits, default_thresholds
from segma_b200.geometry import INFERENCE_SETTINGS, plan_windows
from segma_b200.models import Models

pytestmark = pytest.mark.gpu
LABELS = synth.DEFAULT_LABELS

# fp16 tensor-core operands (11-bit mantissa: 8x tighter than the bf16 tolerance north_star allows) with
# fp32 accumulation / residual stream / LayerNorm / softmax / LSTM state: max |logit error| as a
# fraction of the logit spread of the run
LOGIT_RTOL_OF_STD = 0.01
MIN_LABEL_AGREEMENT = 0.999


def _report(got, ref):
    err = (got - ref).abs()
    std = ref.std().item()
    agree = ((got > 0) == (ref > 0)).float().mean().item()
    return err.max().item(), err.mean().item(), std, agree


#: below this many label decisions one flipped frame is worth more than 0.01 %: the rate cannot be resolved to 99.9 %
#: (2 636 decisions: one flip = 0.038 %), so short fixtures are held to 99.8 % and every model is also run on a
#: fixture of >= 20 000 decisions that is held to the 99.9 % of north_star
RESOLVING_SAMPLE = 10_000
SMALL_SAMPLE_AGREEMENT = 0.998


def _check_logits(got, ref, what, min_agreement=None, max_rtol=LOGIT_RTOL_OF_STD):
    mx, mean, std, agree = _report(got, ref)
    if min_agreement is None:
        min_agreement = MIN_LABEL_AGREEMENT if ref.numel() >= RESOLVING_SAMPLE else SMALL_SAMPLE_AGREEMENT
    print(f"{what}: {ref.numel()} decisions, max|err| {mx:.4g} mean|err| {mean:.4g} logit std {std:.4g} "
          f"label agreement {agree:.5f} (bar {min_agreement})")
    assert mx <= max_rtol * max(std, 1.0), f"{what}: max logit error {mx} vs std {std}"
    assert agree >= min_agreement, f"{what}: frame-label agreement {agree}"


@pytest.mark.parametrize("kind,dims", [("surgical_hydra", synth.WHISPER_TEST), ("hydra_whisper", synth.WHISPER_TEST)])
def test_whisper_file_level(cuda, kind, dims):
    sd = (synth.surgical_hydra_state_dict if kind == "surgical_hydra" else synth.hydra_whisper_state_dict)(dims, seed=3)
    cfg = make_config(kind)
    le = MultiLabelEncoder(list(LABELS))
    model = Models[kind].from_state_dict(sd, le, cfg)
    n = 63680 * 3 + 20000  # 3 full windows (batches of 2 + 1) and a 20000-sample tail
    pcm = synth.synth_audio(n, 11)
    got = apply_model_on_audio(pcm, model, INFERENCE_SETTINGS, "cuda", batch_size=2).cpu()
    fwd = (lambda f: O.surgical_hydra_forward(sd, f, LABELS)) if kind == "surgical_hydra" else (lambda f: O.hydra_whisper_forward(sd, f, LABELS))
    ref = O.apply_model_on_audio(torch.from_numpy(pcm), fwd, 4, batch_size=2, whisper=True)
    assert got.shape == ref.shape == ((n - 400) // 320 + 1, 4)
    _check_logits(got, ref, f"{kind} file-level")
    # decode is bit-exact on identical logits
    thr = default_thresholds(le)
    mask = apply_thresholds(got.cuda(), thr)
    assert torch.equal(mask.cpu(), O.apply_thresholds(got, [0.5] * 4))
    iv = create_intervals(mask, INFERENCE_SETTINGS, le)
    assert iv == O.create_intervals(mask.cpu().numpy(), LABELS)
    assert decode_logits(got.cuda(), thr, le) == iv


def test_whisper_forward_dropin(cuda):
    dims = synth.WHISPER_TEST
    sd = synth.surgical_hydra_state_dict(dims, seed=4)
    le = MultiLabelEncoder(list(LABELS))
    model = Models["surgical_hydra"].from_state_dict(sd, le, make_config("surgical_hydra"))
    wav = [torch.from_numpy(synth.synth_audio(64000, s)) for s in range(3)]
    feats = torch.cat([model.audio_preparation_hook(w) for w in wav])
    assert feats.shape == (3, 80, 3000)
    ref_feats = torch.stack([O.whisper_logmel(w) for w in wav])
    assert (feats.cpu() - ref_feats).abs().max() <= 1e-4
    out = model(feats)
    assert out.shape == (3, 199, 1, 4)
    _check_logits(out.cpu(), O.surgical_hydra_forward(sd, ref_feats, LABELS), "forward drop-in")


def test_whisper_small_dims_few_windows(cuda):
    """BASELINE config 2 model (Whisper-small dims) on a short file."""
    dims = synth.WHISPER_SMALL
    sd = synth.surgical_hydra_state_dict(dims, seed=0)
    le = MultiLabelEncoder(list(LABELS))
    model = Models["surgical_hydra"].from_state_dict(sd, le, make_config("surgical_hydra"))
    n = 63680 * 4 + 320
    pcm = synth.synth_audio(n, 1)
    got = apply_model_on_audio(pcm, model, INFERENCE_SETTINGS, "cuda", batch_size=128).cpu()
    ref = O.apply_model_on_audio(torch.from_numpy(pcm), lambda f: O.surgical_hydra_forward(sd, f, LABELS), 4,
                                 batch_size=128, whisper=True)
    _check_logits(got, ref, "whisper-small dims")


def test_config2_full_size_properties(cuda):
    """BASELINE config 2 at its full size (Whisper-small dims, one 1 h file, batch 128), checked through
    size-independent properties: the frame count of the reference's geometry, the prefix property of the batching
    (the first 128-window batch does not depend on what follows it), and interval decoding that is bit-exact against
    the reference's create_intervals restated on the host boolean mask."""
    dims = synth.WHISPER_SMALL
    sd = synth.surgical_hydra_state_dict(dims, seed=0)
    le = MultiLabelEncoder(list(LABELS))
    model = Models["surgical_hydra"].from_state_dict(sd, le, make_config("surgical_hydra"))
    n = 57_600_000
    pcm = synth.synth_audio(n, 0)
    logits = apply_model_on_audio(pcm, model, INFERENCE_SETTINGS, "cuda", batch_size=128)
    assert logits.shape == (179_999, len(LABELS)) and logits.dtype == torch.float32   # 904 windows + tail
    assert torch.isfinite(logits).all()
    # prefix property: 128 full windows on their own reproduce the first batch bit for bit
    n_prefix = 63_680 * 128 + 320
    first = apply_model_on_audio(pcm[:n_prefix], model, INFERENCE_SETTINGS, "cuda", batch_size=128)
    assert first.shape[0] == 199 * 128
    assert torch.equal(first, logits[: 199 * 128])
    # decode: fused device path == reference semantics on the thresholded mask
    thr = default_thresholds(le)
    mask = apply_thresholds(logits, thr, "cuda")
    got = decode_logits(logits, thr, le)
    ref = O.create_intervals(mask.cpu().numpy(), list(le.base_labels))
    assert got == ref
    assert got == create_intervals(mask, INFERENCE_SETTINGS, le)
    for lab in le.base_labels:  # per label: time-ordered, disjoint, inside the file
        iv = [(s, e) for s, e, l in got if l == lab]
        assert all(0 <= s < e <= 320 * 179_999 for s, e in iv)
        assert all(a[1] < b[0] for a, b in zip(iv, iv[1:]))


def test_overlapping_windows_stitch(cuda):
    dims = synth.WHISPER_TEST
    sd = synth.hydra_whisper_state_dict(dims, seed=5)
    le = MultiLabelEncoder(list(LABELS))
    model = Models["hydra_whisper"].from_state_dict(sd, le, make_config("hydra_whisper"))
    n = 64000 + 32000 * 3 + 5000
    pcm = synth.synth_audio(n, 12)
    step = 32000  # 50 % overlap, multiple of 320
    got = apply_model_on_audio(pcm, model, INFERENCE_SETTINGS, "cuda", batch_size=3, window_step=step).cpu()
    # composed oracle: same windows/batches, logit-domain mean
    n_fit = (n - 64000) // step + 1
    wins, offs = [], []
    t = torch.from_numpy(pcm)
    for b0 in range(0, n_fit, 3):
        idx = list(range(b0, min(b0 + 3, n_fit)))
        feats = torch.stack([O.whisper_logmel(t[i * step: i * step + 64000]) for i in idx])
        out = O.hydra_whisper_forward(sd, feats, LABELS).reshape(len(idx), 199, 4)
        wins += list(out)
        offs += [i * step // 320 for i in idx]
    tail = t[n_fit * step:]
    ft = (tail.numel() - 400) // 320 + 1
    wins.append(O.hydra_whisper_forward(sd, O.whisper_logmel(tail)[None], LABELS).reshape(199, 4)[:ft])
    offs.append(n_fit * step // 320)
    n_frames = max(o + w.shape[0] for o, w in zip(offs, wins))
    ref = O.stitch_mean(wins, offs, n_frames)
    assert got.shape == ref.shape
    _check_logits(got, ref, "50% overlap")


# ---- against the fixtures the reference itself produced (tests/golden/models.npz) ---------------------------
from pathlib import Path  # noqa: E402

GOLDEN = Path(__file__).parent / "golden"


@pytest.mark.parametrize("key,kind,seed,extra", [
    ("surgical_hydra_logits", "surgical_hydra", 3, {}),
    ("surgical_hydra_avg_l2_logits", "surgical_hydra", 8, {"encoder_layers": [2], "reduction": "average"}),
    ("hydra_whisper_logits", "hydra_whisper", 7, {}),
])
def test_whisper_family_vs_reference_golden(cuda, key, kind, seed, extra):
    g = np.load(GOLDEN / "models.npz")
    n, audio_seed, bs = (int(v) for v in g["meta"])
    if kind == "surgical_hydra":
        sd = synth.surgical_hydra_state_dict(synth.WHISPER_TEST, n_mixed_layers=1 if extra else None, seed=seed)
    else:
        sd = synth.hydra_whisper_state_dict(synth.WHISPER_TEST, seed=seed)
    le = MultiLabelEncoder(list(LABELS))
    model = Models[kind].from_state_dict(sd, le, make_config(kind, extra))
    got = apply_model_on_audio(synth.synth_audio(n, audio_seed), model, INFERENCE_SETTINGS, "cuda", batch_size=bs).cpu()
    _check_logits(got, torch.from_numpy(g[key]), f"{key} vs reference golden")


def test_decode_vs_reference_golden(cuda):
    g = np.load(GOLDEN / "models.npz")
    logits = torch.from_numpy(g["hubert_logits"]).cuda()
    le = MultiLabelEncoder(list(LABELS))
    for tname, t in (("t50", 0.5), ("t30", 0.3), ("t70", 0.7)):
        thr = {lab: {"lower_bound": t, "upper_bound": 1.0} for lab in LABELS}
        assert np.array_equal(apply_thresholds(logits, thr).cpu().numpy(), g[f"hubert_mask_{tname}"])
        iv = decode_logits(logits, thr, le)
        want = [(int(s), int(e), LABELS[int(c)]) for c, s, e in g[f"hubert_intervals_{tname}"]]
        assert iv == want
    gi = np.load(GOLDEN / "intervals.npz")
    for k in sorted(k[5:] for k in gi.files if k.startswith("mask_")):
        iv = create_intervals(torch.from_numpy(gi[f"mask_{k}"]), INFERENCE_SETTINGS, le)
        assert iv == [(int(s), int(e), LABELS[int(c)]) for c, s, e in gi[f"iv_{k}"]], k


# ---- wav2vec2 / HuBERT / WavLM family --------------------------------------------------------------------------
@pytest.mark.parametrize("dims,seed", [(synth.W2V2_TEST, 5), (synth.WAVLM_TEST, 6)])
def test_w2v2_family_file_level_small(cuda, dims, seed):
    sd = synth.hubert_hydra_state_dict(dims, seed=seed)
    le = MultiLabelEncoder(list(LABELS))
    model = Models["surgical_hubert_hydra"].from_state_dict(sd, le, make_config("surgical_hubert_hydra"))
    n = 63680 * 3 + 20000
    pcm = synth.synth_audio(n, 11)
    got = apply_model_on_audio(pcm, model, INFERENCE_SETTINGS, "cuda", batch_size=2).cpu()
    ref = O.apply_model_on_audio(torch.from_numpy(pcm), lambda w: O.hubert_hydra_forward(sd, w, LABELS), 4, batch_size=2)
    assert got.shape == ref.shape == ((n - 400) // 320 + 1, 4)
    _check_logits(got, ref, f"w2v2 wavlm={dims.wavlm} file-level")
    # forward drop-in on (B, n_samples)
    wav = torch.stack([torch.from_numpy(synth.synth_audio(64000, s)) for s in range(2)])
    out = model(wav)
    assert out.shape == (2, 199, 1, 4)
    _check_logits(out.cpu(), O.hubert_hydra_forward(sd, wav, LABELS), "w2v2 forward drop-in")


@pytest.mark.parametrize("key,dims,seed", [("hubert_logits", synth.HUBERT_BASE, 5), ("wavlm_logits", synth.WAVLM_BASE, 6)])
def test_w2v2_family_vs_reference_golden(cuda, key, dims, seed):
    """BASELINE configs 1 and 3 (HuBERT-base / WavLM-base+ dims) against logits the reference's own
    apply_model_on_audio produced (tests/golden/models.npz)."""
    g = np.load(GOLDEN / "models.npz")
    n, audio_seed, bs = (int(v) for v in g["meta"])
    sd = synth.hubert_hydra_state_dict(dims, seed=seed)
    le = MultiLabelEncoder(list(LABELS))
    model = Models["surgical_hubert_hydra"].from_state_dict(sd, le, make_config("surgical_hubert_hydra"))
    got = apply_model_on_audio(synth.synth_audio(n, audio_seed), model, INFERENCE_SETTINGS, "cuda", batch_size=bs).cpu()
    _check_logits(got, torch.from_numpy(g[key]), f"{key} vs reference golden")


@pytest.mark.parametrize("key,dims,seed", [("hubert_logits", synth.HUBERT_BASE, 5), ("wavlm_logits", synth.WAVLM_BASE, 6)])
def test_w2v2_family_vs_reference_golden_large(cuda, key, dims, seed):
    """The 99.9 % frame-label bar of north_star on a sample that can resolve it: 26 windows + tail = 20 800 label
    decisions per model, logits from the reference's own ``apply_model_on_audio`` (tests/golden/models_large.npz,
    oracle/make_golden.py::models_large; hubert/surgical_hydra.py:87-101)."""
    g = np.load(GOLDEN / "models_large.npz")
    n, audio_seed, bs = (int(v) for v in g["meta"])
    ref = torch.from_numpy(g[key])
    assert ref.numel() >= 20_000
    sd = synth.hubert_hydra_state_dict(dims, seed=seed)
    le = MultiLabelEncoder(list(LABELS))
    model = Models["surgical_hubert_hydra"].from_state_dict(sd, le, make_config("surgical_hubert_hydra"))
    got = apply_model_on_audio(synth.synth_audio(n, audio_seed), model, INFERENCE_SETTINGS, "cuda", batch_size=bs).cpu()
    assert got.shape == ref.shape
    _check_logits(got, ref, f"{key} (26 windows) vs reference golden")


def test_whisper_config2_full_batch_vs_reference_golden(cuda):
    """BASELINE config 2's model on one full 128-window forward call + a remainder batch of 8 + the tail: the LSTM
    recurrence runs over all 128 windows (whisper/surgical_hydra.py:57-60,101; batch boundaries of
    inference.py:138-206).  Logits come from the reference's ``SurgicalHydra`` class at Whisper-small dims
    (tests/golden/whisper_config2.npz, oracle/make_golden.py::whisper_config2); every logit is compared."""
    g = np.load(GOLDEN / "whisper_config2.npz")
    n, audio_seed, bs = (int(v) for v in g["meta"])
    ref = torch.from_numpy(g["logits"])
    assert bs == 128 and ref.shape == (136 * 199 + 103, 4)
    sd = synth.surgical_hydra_state_dict(synth.WHISPER_SMALL, seed=0)
    le = MultiLabelEncoder(list(LABELS))
    model = Models["surgical_hydra"].from_state_dict(sd, le, make_config("surgical_hydra"))
    got = apply_model_on_audio(synth.synth_audio(n, audio_seed), model, INFERENCE_SETTINGS, "cuda", batch_size=bs).cpu()
    assert got.shape == ref.shape
    # Stated tolerance for the 128-step recurrence (profiles/r02c_diag_whisper_recurrence.txt):
    # the windows of a file resemble each other (the component common to all windows is 6x the varying one), so the
    # reference's own LSTM integrates any *systematic* perturbation of its input along the window axis -- in exact
    # fp32 arithmetic the 5.6e-4 relative error that fp16 weight rounding leaves in the encoder output, identical for
    # every window, moves single logits of a few frame rows by up to 0.1 (11 % of the spread) while the 4.5e-4
    # window-varying part moves none by more than 0.003.  The LSTM path itself is exact to 2e-4 (split-precision
    # projection, fp32-level W_hh).  Hence: >= 99.9 % label agreement (north_star), mean error <= 0.25 %, 99 % of
    # the logits within the 1 % bar of the short-recurrence tests, every logit within 15 % of the spread.
    err = (got - ref).abs()
    std = ref.std().item()
    _check_logits(got, ref, "whisper-small dims, 128 + 8 windows + tail vs reference golden", max_rtol=0.15)
    q99 = torch.quantile(err.reshape(-1).double(), 0.99).item()
    print(f"mean|err|/std {err.mean().item() / std:.4%}  q99|err|/std {q99 / std:.4%}")
    assert err.mean().item() <= 0.0025 * std
    assert q99 <= LOGIT_RTOL_OF_STD * std
    # the remainder batch (8 windows) and the tail are short recurrences: the 1 % bar holds for every logit
    _check_logits(got[128 * 199:], ref[128 * 199:], "remainder batch + tail")


def test_whisper_encoder_output_error(cuda):
    """What the LSTM consumes: the layer-weighted encoder output ``mix`` of Whisper-small dims against the oracle's
    hidden states (4 windows).  fp16 tensor-core operands leave <= 1.5e-3 of its rms (measured 7e-4); the LSTM tail
    behind it is checked to fp32 level in test_kernels_gpu.py::test_lstm_stack_split_precision_projection."""
    sd = synth.surgical_hydra_state_dict(synth.WHISPER_SMALL, seed=0)
    le = MultiLabelEncoder(list(LABELS))
    model = Models["surgical_hydra"].from_state_dict(sd, le, make_config("surgical_hydra"))
    wav = [torch.from_numpy(synth.synth_audio(64000, 70 + s)) for s in range(4)]
    feats = torch.stack([O.whisper_logmel(w) for w in wav])
    model(feats.cuda())
    got = model.engine._ws["mix"][:4].float().cpu()
    with torch.inference_mode():
        hs = O.whisper_encoder_hidden_states(sd, feats)[1:]
        w = torch.softmax(sd["layer_weights"], 0)
        ref = sum(w[j] * hs[j][:, :199] for j in range(len(hs)))
    rms = ref.pow(2).mean().sqrt().item()
    e = (got - ref)
    print(f"mix rms {rms:.4g}: rms err / rms {e.pow(2).mean().sqrt().item() / rms:.3e}, max err / rms {e.abs().max().item() / rms:.3e}")
    assert e.pow(2).mean().sqrt().item() <= 1.5e-3 * rms and e.abs().max().item() <= 1.5e-2 * rms


def test_w2v2_silent_file(cuda):
    """The reference's io fixture (all-zero audio): GroupNorm variance 0 everywhere; logits are finite and constant in time."""
    sd = synth.hubert_hydra_state_dict(synth.W2V2_TEST, seed=5)
    le = MultiLabelEncoder(list(LABELS))
    model = Models["surgical_hubert_hydra"].from_state_dict(sd, le, make_config("surgical_hubert_hydra"))
    pcm = np.zeros(64000 + 63680, dtype=np.float32)
    got = apply_model_on_audio(pcm, model, INFERENCE_SETTINGS, "cuda", batch_size=4).cpu()
    ref = O.apply_model_on_audio(torch.from_numpy(pcm), lambda w: O.hubert_hydra_forward(sd, w, LABELS), 4, batch_size=4)
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max() <= 0.01 * max(1.0, ref.abs().max().item())


# ---- BASELINE config 5: window length / overlap / threshold sweep -------------------------------------------------
def _composed_oracle(pcm, forward, win, step, batch_size, whisper, frames_per_window):
    """Same windows and batch boundaries as the product, logit-domain mean over covering windows."""
    t = torch.from_numpy(pcm)
    n = t.numel()
    n_fit = (n - win) // step + 1 if n >= win else 0
    wins, offs = [], []
    for b0 in range(0, n_fit, batch_size):
        idx = list(range(b0, min(b0 + batch_size, n_fit)))
        chunk = [t[i * step: i * step + win] for i in idx]
        x = torch.stack([O.whisper_logmel(w) for w in chunk]) if whisper else torch.stack(chunk)
        out = forward(x).reshape(len(idx), -1, len(LABELS))[:, :frames_per_window]
        wins += list(out)
        offs += [i * step // 320 for i in idx]
    tail = t[n_fit * step:]
    if tail.numel() >= 400:
        ft = min((tail.numel() - 400) // 320 + 1, frames_per_window)
        x = O.whisper_logmel(tail)[None] if whisper else tail[None]
        wins.append(forward(x).reshape(-1, len(LABELS))[:ft])
        offs.append(n_fit * step // 320)
    n_frames = max(o + w.shape[0] for o, w in zip(offs, wins))
    return O.stitch_mean(wins, offs, n_frames)


# BASELINE config 5: the full window x overlap grid (thresholds 0.3 / 0.5 / 0.7 are swept inside the test)
@pytest.mark.parametrize("win_s", [2, 3, 4, 6, 8])
@pytest.mark.parametrize("overlap", [0.5, 0.75, 0.9])
def test_sweep_waveform_model(cuda, win_s, overlap):
    sd = synth.hubert_hydra_state_dict(synth.W2V2_TEST, seed=5)
    le = MultiLabelEncoder(list(LABELS))
    cfg = make_config("surgical_hubert_hydra", chunk_duration_s=float(win_s))
    model = Models["surgical_hubert_hydra"].from_state_dict(sd, le, cfg)
    win = win_s * 16000
    step = max(320, int(win * (1 - overlap)) // 320 * 320)
    F_ = (win - 400) // 320 + 1
    n = win + step * 5 + 4000
    pcm = synth.synth_audio(n, 40 + win_s)
    got = apply_model_on_audio(pcm, model, INFERENCE_SETTINGS, "cuda", batch_size=4, chunk_duration_s=float(win_s),
                               window_step=step)
    ref = _composed_oracle(pcm, lambda w: O.hubert_hydra_forward(sd, w, LABELS), win, step, 4, False, F_)
    assert got.shape == ref.shape
    _check_logits(got.cpu(), ref, f"sweep {win_s}s overlap {overlap}")
    for t in (0.3, 0.5, 0.7):  # decoding of the product's own logits is bit-exact at every threshold
        thr = {lab: {"lower_bound": t, "upper_bound": 1.0} for lab in LABELS}
        mask = O.apply_thresholds(got.cpu(), [t] * 4)
        assert decode_logits(got, thr, le) == O.create_intervals(mask.numpy(), LABELS)


@pytest.mark.parametrize("win_s", [2, 6])
def test_sweep_whisper_window_length(cuda, win_s):
    """Whisper-family windows other than 4 s keep conv_settings.n_windows(chunk, strict=False) frames."""
    sd = synth.hydra_whisper_state_dict(synth.WHISPER_TEST, seed=7)
    le = MultiLabelEncoder(list(LABELS))
    cfg = make_config("hydra_whisper", chunk_duration_s=float(win_s))
    model = Models["hydra_whisper"].from_state_dict(sd, le, cfg)
    win = win_s * 16000
    F_ = win // 321
    assert model.n_keep == F_
    step = F_ * 320  # windows tile the frame grid
    n = win + step * 2 + 3000
    pcm = synth.synth_audio(n, 50 + win_s)
    got = apply_model_on_audio(pcm, model, INFERENCE_SETTINGS, "cuda", batch_size=2, chunk_duration_s=float(win_s),
                               window_step=step)
    ref = _composed_oracle(pcm, lambda f: O.hydra_whisper_forward(sd, f, LABELS, n_keep=F_), win, step, 2, True, F_)
    assert got.shape == ref.shape
    _check_logits(got.cpu(), ref, f"whisper {win_s}s windows")


# ---- audio decode + staging; f2: RTTM / logits artefacts -----------------------------------------
@pytest.mark.parametrize("subtype", ["int16", "float32"])
def test_wav_staging_matches_host_decode(cuda, tmp_path, subtype):
    from segma_b200.io import get_all_samples, stage_to_device, write_wav

    pcm = synth.synth_audio(100_003, 9)
    p = tmp_path / "a.wav"
    write_wav(p, pcm, subtype=subtype)
    dev = stage_to_device(p, "cuda")
    assert torch.equal(dev.cpu(), get_all_samples(p)[0])  # bit-identical to the host decoder


def test_infer_file_on_wav_writes_reference_artefacts(cuda, tmp_path):
    from segma_b200.inference import infer_file
    from segma_b200.io import write_wav

    sd = synth.hubert_hydra_state_dict(synth.W2V2_TEST, seed=5)
    le = MultiLabelEncoder(list(LABELS))
    cfg = make_config("surgical_hubert_hydra")
    model = Models["surgical_hubert_hydra"].from_state_dict(sd, le, cfg)
    pcm = synth.synth_audio(64000 + 63680 + 7000, 13)
    wav = tmp_path / "rec_01.wav"
    write_wav(wav, pcm, subtype="int16")
    intervals = infer_file(wav, model, tmp_path / "out", cfg, batch_size=2, device="cuda", save_logits=True)
    # logits artefact in the format scripts/tune.py:95-113 reads: {label: (n_frames,) fp32}
    blob = torch.load(tmp_path / "out" / "logits" / "rec_01-logits_dict_t.pt")
    assert list(blob) == list(LABELS)
    logits = torch.stack([blob[lab] for lab in LABELS], dim=1)
    assert logits.shape == ((pcm.size - 400) // 320 + 1, 4) and logits.dtype == torch.float32
    # RTTM text is the reference's format and decodes the saved logits bit-exactly
    want = O.create_intervals(O.apply_thresholds(logits, [0.5] * 4).numpy(), LABELS)
    assert intervals == want
    lines = (tmp_path / "out" / "raw_rttm" / "rec_01.rttm").read_text().splitlines()
    assert len(lines) == len(want)
    for line, (s, e, lab) in zip(lines, want):
        assert line == f"SPEAKER rec_01 <NA> {round(s / 16000, 8)} {round((e - s) / 16000, 8)} <NA> <NA> {lab} <NA> <NA>"
    # the int16 file was decoded exactly like the host path would
    pcm16 = np.clip(np.round(pcm * 32768.0), -32768, 32767).astype(np.int16).astype(np.float32) / 32768.0
    ref = O.apply_model_on_audio(torch.from_numpy(pcm16), lambda w: O.hubert_hydra_forward(sd, w, LABELS), 4, batch_size=2)
    _check_logits(logits, ref, "infer_file on int16 wav")


def test_run_inference_on_audios_cli_path(cuda, tmp_path, capsys):
    """The whole driver as the reference's __main__ calls it: YAML config, Lightning-style checkpoint, a folder
    of wav files, a thresholds YAML -> one RTTM per file, identical to decoding the oracle-checked logits."""
    import yaml

    from segma_b200.inference import run_inference_on_audios
    from segma_b200.io import write_wav

    sd = synth.hubert_hydra_state_dict(synth.W2V2_TEST, seed=5)
    torch.save({"state_dict": sd, "epoch": 3}, tmp_path / "best.ckpt")
    cfg = make_config("surgical_hubert_hydra")
    cfg.save(tmp_path / "config.yml")
    wavs = tmp_path / "wav"
    wavs.mkdir()
    lens = {"b_second": 64000 + 9000, "a_first": 3 * 63680 + 500, "c_short": 300}
    for name, n in lens.items():
        write_wav(wavs / f"{name}.wav", synth.synth_audio(n, len(name)), subtype="float32")
    thr = {lab: {"lower_bound": 0.4, "upper_bound": 1.0} for lab in LABELS}
    (tmp_path / "thr.yml").write_text(yaml.safe_dump(thr))
    done = run_inference_on_audios(config=tmp_path / "config.yml", uris=None, wavs=wavs, checkpoint=tmp_path / "best.ckpt",
                                   output=tmp_path / "out", thresholds=tmp_path / "thr.yml", batch_size=2, device="gpu",
                                   save_logits=True)
    assert [p.stem for p in done] == ["a_first", "b_second", "c_short"]  # sorted, like the reference
    log = capsys.readouterr().out
    assert "[log] - (1/3) - running inference for file: 'a_first'" in log
    for name, n in lens.items():
        rttm = (tmp_path / "out" / "raw_rttm" / f"{name}.rttm").read_text().splitlines()
        if n < 400:
            assert rttm == []  # shorter than one receptive field: no frames, empty RTTM
            continue
        blob = torch.load(tmp_path / "out" / "logits" / f"{name}-logits_dict_t.pt")
        logits = torch.stack([blob[lab] for lab in LABELS], dim=1)
        assert logits.shape[0] == (n - 400) // 320 + 1
        want = O.create_intervals(O.apply_thresholds(logits, [0.4] * 4).numpy(), LABELS)
        assert len(rttm) == len(want)
        assert rttm[:3] == [f"SPEAKER {name} <NA> {round(s / 16000, 8)} {round((e - s) / 16000, 8)} <NA> <NA> {lab} <NA> <NA>"
                            for s, e, lab in want[:3]]


def test_packed_windows_of_several_files_equal_file_by_file(cuda):
    """``apply_model_on_audios`` packs the (independent) windows of wav2vec2-family models across file boundaries into
    full forward calls and lets tails of equal length share a call: the logits are the same bits as file-by-file
    ``apply_model_on_audio``, and ``infer_corpus`` returns the same table as decoding each file on its own."""
    from segma_b200.inference import apply_model_on_audios, infer_corpus

    sd = synth.hubert_hydra_state_dict(synth.W2V2_TEST, seed=5)
    le = MultiLabelEncoder(list(LABELS))
    cfg = make_config("surgical_hubert_hydra")
    model = Models["surgical_hubert_hydra"].from_state_dict(sd, le, cfg)
    lens = [64000 + 63680 * 2 + 9000, 300, 64000, 70_000, 64000 + 63680 + 9000, 5000, 63680 * 4 + 320 + 70_000 - 64000, 0, 12_345]
    files = [synth.synth_audio(n, 60 + i) if n else np.zeros(0, dtype=np.float32) for i, n in enumerate(lens)]
    packed = apply_model_on_audios(files, model, INFERENCE_SETTINGS, "cuda", batch_size=3)
    assert len(packed) == len(files)
    for f, got in zip(files, packed):
        want = apply_model_on_audio(f, model, INFERENCE_SETTINGS, "cuda", batch_size=3)
        assert got.shape == want.shape == (max((f.size - 400) // 320 + 1, 0) if f.size >= 400 else 0, 4)
        assert torch.equal(got, want)
    table = infer_corpus(files, model, cfg, batch_size=3, device="cuda").cpu().numpy()
    thr = default_thresholds(le)
    rows = []
    for i, f in enumerate(files):
        iv = decode_logits(apply_model_on_audio(f, model, INFERENCE_SETTINGS, "cuda", batch_size=3), thr, le)
        rows += [(i, LABELS.index(lab), s, e) for s, e, lab in iv]
    assert table.tolist() == [list(r) for r in rows]
    # Whisper-family corpus: short files take turns on several streams (infer_corpus), same table as one by one
    sdw = synth.hydra_whisper_state_dict(synth.WHISPER_TEST, seed=7)
    cfgw = make_config("hydra_whisper")
    mw = Models["hydra_whisper"].from_state_dict(sdw, le, cfgw)
    small = [synth.synth_audio(n, 90 + i) for i, n in enumerate([64000 + 5000, 70_000, 63680 * 2 + 64000, 9000, 64000, 63680 + 64000 + 700, 30_000])]
    tw = infer_corpus(small, mw, cfgw, batch_size=2, device="cuda").cpu().numpy()
    rows = []
    for i, f in enumerate(small):
        iv = decode_logits(apply_model_on_audio(f, mw, INFERENCE_SETTINGS, "cuda", batch_size=2), thr, le)
        rows += [(i, LABELS.index(lab), s, e) for s, e, lab in iv]
    assert tw.tolist() == [list(r) for r in rows]
    # Whisper-family models are never packed (the LSTM couples the windows of a call): one file after the other
    two = [synth.synth_audio(64000 + 63680 + 3000, 80), synth.synth_audio(64000, 81)]
    for f, got in zip(two, apply_model_on_audios(two, mw, INFERENCE_SETTINGS, "cuda", batch_size=2)):
        assert torch.equal(got, apply_model_on_audio(f, mw, INFERENCE_SETTINGS, "cuda", batch_size=2))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["hydra_whisper", "surgical_hubert_hydra"])
def test_corpus_partitioned_by_file_and_window_batch(cuda, kind):
    """Multi-GPU partition by audio file *and* window batch, emulated rank by rank on one GPU: a long file
    is cut into batch ranges, each rank decodes its units on their own, and the shifted + merged pieces equal the table
    of a single process bit for bit; the logits of a batch range are the same bits as that slice of the whole file."""
    from segma_b200.distributed import merge_split_files
    from segma_b200.geometry import assign_units, batch_frame_range, plan_work_units
    from segma_b200.inference import infer_corpus

    le = MultiLabelEncoder(list(LABELS))
    cfg = make_config(kind)
    if kind == "hydra_whisper":
        model = Models[kind].from_state_dict(synth.hydra_whisper_state_dict(synth.WHISPER_TEST, seed=7), le, cfg)
        fpw = model.n_keep
    else:
        model = Models[kind].from_state_dict(synth.hubert_hydra_state_dict(synth.W2V2_TEST, seed=5), le, cfg)
        fpw = 199
    bs = 2
    lens = [63680 * 21 + 64000 + 7000, 64000, 300, 63680 * 2 + 64000, 70_000, 63680 * 9 + 320 * 30, 12_345]
    files = [synth.synth_audio(n, 40 + i) for i, n in enumerate(lens)]
    # make activity cross the cuts: thresholds low enough that long runs exist
    thr = {lab: {"lower_bound": 0.35, "upper_bound": 1.0} for lab in LABELS}
    whole = infer_corpus(files, model, cfg, batch_size=bs, device="cuda", thresholds=thr).cpu()
    for world in (2, 4):
        units = plan_work_units(lens, world, 64000, bs, 63680, fpw)
        assert any(not u.whole_file for u in units), "the long file must be cut for this test to mean anything"
        parts = [infer_corpus(files, model, cfg, batch_size=bs, device="cuda", thresholds=thr, shard=(r, world), gather=False)
                 for r in range(world)]
        assert sum(p.shape[0] for p in parts) >= whole.shape[0]
        merged = merge_split_files(torch.cat(parts)).cpu()
        assert torch.equal(merged, whole), (world, merged.shape, whole.shape)
    # the logits of one unit are that slice of the whole file's logits
    full = apply_model_on_audio(files[0], model, INFERENCE_SETTINGS, "cuda", batch_size=bs)
    plan = plan_windows(lens[0], 64000, bs, 63680, fpw)
    for lo, hi in ((0, 3), (3, 4), (len(plan.batches) - 2, len(plan.batches))):
        f_lo, f_hi = batch_frame_range(plan, lo, hi)
        got = apply_model_on_audio(files[0], model, INFERENCE_SETTINGS, "cuda", batch_size=bs, batch_range=(lo, hi))
        assert got.shape[0] == f_hi - f_lo and torch.equal(got, full[f_lo:f_hi])
    assert assign_units(units, 4) == assign_units(units, 4)


# ---- the reference's own forward tests (tests/test_models.py:37-71), widened over dims, labels and LSTM shapes ------------
@pytest.mark.parametrize("kind,dims,lstm,labels", [
    ("surgical_hydra", synth.WHISPER_TEST, synth.LSTMDims(), ("aa", "bb", "cc")),            # the reference's 3 labels
    ("hydra_whisper", synth.WHISPER_TEST, synth.LSTMDims(), ("aa", "bb", "cc")),
    ("surgical_hydra", synth.WHISPER_TINY, synth.LSTMDims(64, 1, False), ("only",)),          # d=384: 128-wide N tiles; 1 label
    ("hydra_whisper", synth.WHISPER_BASE, synth.LSTMDims(256, 1, True), tuple("abcdefg")),    # d=512; H=256 (L2-resident W_hh)
    ("surgical_hydra", synth.WHISPER_TEST, synth.LSTMDims(128, 3, True), tuple("abcdefghijkl")),  # 3 LSTM layers, 12 labels
])
def test_whisper_based_forward_like_the_reference_tests(cuda, kind, dims, lstm, labels):
    """``model(x_t)`` on ``torch.ones((2, 80, 3000))`` as /root/reference/tests/test_models.py:37-53 calls it (the
    reference only checks that it runs; here the logits are compared with the oracle), over encoder widths that take
    different GEMM tile shapes, LSTM shapes that take different recurrence kernels, and 1 ... 12 labels (decode kernels
    for C <= 8 and C > 8)."""
    make = synth.surgical_hydra_state_dict if kind == "surgical_hydra" else synth.hydra_whisper_state_dict
    sd = make(dims, lstm=lstm, labels=labels, seed=21)
    le = MultiLabelEncoder(list(labels))
    cfg = make_config(kind, {"lstm": {"hidden_size": lstm.hidden_size, "num_layers": lstm.num_layers,
                                      "bidirectional": lstm.bidirectional, "dropout": 0.5}}, classes=labels)
    model = Models[kind].from_state_dict(sd, le, cfg)
    x_t = torch.ones((2, 80, 3000))
    out = model(x_t)
    fwd = O.surgical_hydra_forward if kind == "surgical_hydra" else O.hydra_whisper_forward
    ref = fwd(sd, x_t, labels)
    if isinstance(out, dict):  # HydraWhisper returns the per-head dict (hydra.py:83-87)
        assert list(out) == [f"linear_head_{lab}" for lab in labels]
        out = torch.stack([out[f"linear_head_{lab}"] for lab in labels], dim=-1)
    assert out.shape == ref.shape == (2, 199, 1, len(labels))
    _check_logits(out.cpu().reshape(-1, len(labels)), ref.reshape(-1, len(labels)), f"{kind} d={dims.d_model} H={lstm.hidden_size}")
    # a real signal through the same model, file level, decoded: intervals bit-exact on the product's own logits
    pcm = synth.synth_audio(64000 + 63680 + 9000, 33)
    logits = apply_model_on_audio(pcm, model, INFERENCE_SETTINGS, "cuda", batch_size=2)
    thr = default_thresholds(le)
    mask = O.apply_thresholds(logits.cpu(), [0.5] * len(labels))
    assert decode_logits(logits, thr, le) == O.create_intervals(mask.numpy(), list(le.base_labels))


def test_wavlm_based_forward_like_the_reference_tests(cuda):
    """``model(torch.ones((2, 32_000)))`` on raw audio (/root/reference/tests/test_models.py:56-71), 3 labels; constant
    audio makes GroupNorm's variance zero in layer 0 (the io fixture's degenerate case)."""
    labels = ("aa", "bb", "cc")
    for dims, seed in ((synth.W2V2_TEST, 22), (synth.WAVLM_TEST, 23)):
        sd = synth.hubert_hydra_state_dict(dims, labels=labels, seed=seed)
        model = Models["surgical_hubert_hydra"].from_state_dict(sd, MultiLabelEncoder(list(labels)),
                                                                make_config("surgical_hubert_hydra", classes=labels))
        x_t = torch.ones((2, 32_000))
        out = model(x_t)
        ref = O.hubert_hydra_forward(sd, x_t, labels)
        assert out.shape == ref.shape == (2, 99, 1, 3)
        assert torch.isfinite(out).all()
        assert (out.cpu() - ref).abs().max() <= 0.01 * max(1.0, ref.abs().max().item())


# ---- BASELINE configs 1 and 3 at their stated sizes ---------------------------------------------------------------------
@pytest.mark.parametrize("which", ["hubert", "whisper_tiny"])
def test_config1_180s_file_vs_oracle(cuda, which):
    """BASELINE config 1: the smallest models on a 180 s file = 2 880 000 samples -> 45 windows in one
    remainder batch + a 14 400-sample tail = 8 999 frames, against the CPU oracle on every logit (35 996 decisions: the
    99.9 % bar resolves)."""
    n = 2_880_000
    pcm = synth.synth_audio(n, 5)
    le = MultiLabelEncoder(list(LABELS))
    if which == "hubert":
        sd = synth.hubert_hydra_state_dict(synth.HUBERT_BASE, seed=5)
        model = Models["surgical_hubert_hydra"].from_state_dict(sd, le, make_config("surgical_hubert_hydra"))
        fwd, whisper = (lambda w: O.hubert_hydra_forward(sd, w, LABELS)), False
    else:
        sd = synth.surgical_hydra_state_dict(synth.WHISPER_TINY, seed=6)
        model = Models["surgical_hydra"].from_state_dict(sd, le, make_config("surgical_hydra"))
        fwd, whisper = (lambda f: O.surgical_hydra_forward(sd, f, LABELS)), True
    got = apply_model_on_audio(pcm, model, INFERENCE_SETTINGS, "cuda", batch_size=128).cpu()
    assert got.shape == (8_999, 4)
    torch.set_num_threads(16)
    ref = O.apply_model_on_audio(torch.from_numpy(pcm), fwd, 4, batch_size=128, whisper=whisper)
    _check_logits(got, ref, f"config 1, {which}, 180 s file")
    thr = default_thresholds(le)
    assert decode_logits(got.cuda(), thr, le) == O.create_intervals(O.apply_thresholds(got, [0.5] * 4).numpy(), LABELS)


def test_config3_full_size_properties(cuda):
    """BASELINE config 3 at size: WavLM-base+ dims on hours of audio (two 1 h files and a few short ones through the
    corpus driver), checked through size-independent properties: the reference's frame count per file, independence of
    the files (the packed corpus path gives each file the logits it gets on its own), and interval decoding that is
    bit-exact against the reference's create_intervals restated on the thresholded mask."""
    from segma_b200.inference import apply_model_on_audios, infer_corpus

    sd = synth.hubert_hydra_state_dict(synth.WAVLM_BASE, seed=6)
    le = MultiLabelEncoder(list(LABELS))
    cfg = make_config("surgical_hubert_hydra")
    model = Models["surgical_hubert_hydra"].from_state_dict(sd, le, cfg)
    hour = synth.synth_audio(57_600_000, 0)
    files = [hour, hour[1_000_000:1_000_000 + 700_000], hour[::-1].copy(), hour[5_000_000:5_064_000], hour[:33_280]]
    per_file = apply_model_on_audios(files, model, INFERENCE_SETTINGS, "cuda", batch_size=128)
    assert [t.shape[0] for t in per_file] == [(f.size - 400) // 320 + 1 for f in files]
    assert per_file[0].shape == (179_999, 4) and all(torch.isfinite(t).all() for t in per_file)
    for k in (1, 3, 4):  # short files: the same bits as on their own
        assert torch.equal(per_file[k], apply_model_on_audio(files[k], model, INFERENCE_SETTINGS, "cuda", batch_size=128))
    # a window's logits do not depend on its neighbours: the first 10 windows of the hour alone
    head = apply_model_on_audio(hour[: 63_680 * 10 + 320], model, INFERENCE_SETTINGS, "cuda", batch_size=128)
    assert torch.equal(head, per_file[0][: head.shape[0]])
    thr = default_thresholds(le)
    table = infer_corpus(files, model, cfg, batch_size=128, device="cuda").cpu().numpy()
    rows = []
    for i, t in enumerate(per_file):
        mask = O.apply_thresholds(t.cpu(), [0.5] * 4)
        rows += [(i, LABELS.index(lab), s, e) for s, e, lab in O.create_intervals(mask.numpy(), LABELS)]
    assert table.shape[0] == len(rows) and table.tolist() == [list(r) for r in rows]


def test_second_device_in_one_process(cuda):
    """A process may drive more than one GPU: ``model.to("cuda:1")`` moves the packed weights, every entry point runs
    on the device it is given whatever the current device is, and the library's per-device state (function attributes,
    mel tables, SM count) is set up per device.  Same logits on both devices, bit for bit."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    le = MultiLabelEncoder(list(LABELS))
    pcm = synth.synth_audio(64000 + 63680 * 2 + 9000, 17)
    thr = default_thresholds(le)
    for kind, sd in (("surgical_hydra", synth.surgical_hydra_state_dict(synth.WHISPER_TEST, seed=3)),
                     ("surgical_hubert_hydra", synth.hubert_hydra_state_dict(synth.WAVLM_TEST, seed=6))):
        model = Models[kind].from_state_dict(sd, le, make_config(kind))
        a = apply_model_on_audio(pcm, model, INFERENCE_SETTINGS, "cuda:0", batch_size=2)
        iv0 = decode_logits(a, thr, le)
        model.to("cuda:1")
        assert torch.cuda.current_device() == 0
        b = apply_model_on_audio(pcm, model, INFERENCE_SETTINGS, "cuda:1", batch_size=2)
        assert b.device == torch.device("cuda", 1) and torch.cuda.current_device() == 0
        assert torch.equal(a.cpu(), b.cpu())
        assert decode_logits(b, thr, le) == iv0
        with pytest.raises(Exception):  # weights on cuda:1, asked for cuda:0: refused, not silently wrong
            apply_model_on_audio(pcm, model, INFERENCE_SETTINGS, "cuda:0", batch_size=2)
