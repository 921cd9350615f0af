"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the REFERENCE itself
(/root/reference, imported through oracle/ref_shim.py) on seeded synthetic inputs.

    python oracle/make_golden.py [geometry logmel intervals models models_large whisper_config2 postprocess tuning]

Inputs are not stored: every fixture records the seeds, and tests regenerate audio and weights with
``segma_b200.synth``.  Outputs are the reference's own: ``ConvolutionSettings`` values, the Whisper
feature extractor's log-mel, ``Models[...]`` forwards, ``apply_model_on_audio`` / ``apply_thresholds`` /
``create_intervals`` of src/segma/inference.py.  Only this container has the reference; the fixtures
travel to the GPU box.
"""
from __future__ import annotations

import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import ref_shim  # noqa: E402

ref_shim.install()
import segma.inference as inf  # noqa: E402
from segma.config.base import HydraWhisperConfig, LSTMConfig, SurgicalHydraConfig, SurgicalHydraLightHuBERTConfig  # noqa: E402
from segma.models import Models  # noqa: E402
from segma.models.base import ConvolutionSettings  # noqa: E402
from segma.utils.encoders import MultiLabelEncoder  # noqa: E402
from transformers import WhisperConfig, WhisperFeatureExtractor  # noqa: E402
from transformers.models.whisper.modeling_whisper import WhisperEncoder  # noqa: E402

from segma_b200 import synth  # noqa: E402

OUT = ROOT / "tests" / "golden"
LABELS = synth.DEFAULT_LABELS
torch.set_num_threads(8)


def geometry():
    rows = []
    for ks, ss, ps in [((3, 2), (3, 1), (1, 0)), ((2,), (1,), (0,)), ((320,), (320,), (0,)), ((400, 3, 3), (160, 1, 2), (200, 1, 1)),
                       ((10, 3, 3, 3, 3, 2, 2), (5, 2, 2, 2, 2, 2, 2), (0,) * 7)]:
        cs = ConvolutionSettings(ks, ss, ps)
        for u in (0, 1, 7, 198):
            rows.append((len(ks), *ks, *ss, *ps, u, cs.rf_start_i(u), cs.rf_end_i(u), cs.rf_size, cs.rf_step))
    nw = []
    for chunk in (32000, 48000, 64000, 96000, 112000, 128000):
        icfg = ConvolutionSettings((320,), (320,), (0,))
        wcfg = ConvolutionSettings((400, 3, 3), (160, 1, 2), (200, 1, 1))
        nw.append((chunk, icfg.n_windows(chunk, True), wcfg.n_windows(chunk, False)))
    # frame counts of apply_model_on_audio come from the model fixtures below
    np.savez(OUT / "geometry.npz", rf=np.array([np.array(r + (0,) * (30 - len(r))) for r in rows]), n_windows=np.array(nw))


def logmel():
    fe = WhisperFeatureExtractor()
    out = {}
    for name, n, seed in [("win", 64000, 1), ("tail", 33280, 2), ("short", 1000, 3)]:
        x = synth.synth_audio(n, seed)
        f = fe(x, return_tensors="pt", sampling_rate=16000)["input_features"][0].numpy()
        nv = -(-(n + 200) // 160)
        out[f"{name}_valid"] = f[:, :nv + 2]
        out[f"{name}_fill"] = f[:, nv + 2:].copy()[:, :1]
        assert np.all(f[:, nv:] == f[:, nv:nv + 1]), "frames beyond the audio must be constant"
        out[f"{name}_meta"] = np.array([n, seed, nv])
    z = fe(np.zeros(64000, dtype=np.float32), return_tensors="pt", sampling_rate=16000)["input_features"][0].numpy()
    out["zeros_value"] = np.array([z.min(), z.max()])
    out["mel_filters"] = fe.mel_filters
    np.savez(OUT / "logmel.npz", **out)


def _whisper_dir(dims):
    cfgw = WhisperConfig(d_model=dims.d_model, encoder_layers=dims.n_layers, encoder_attention_heads=dims.n_heads,
                         encoder_ffn_dim=dims.ffn, decoder_layers=1, decoder_attention_heads=2, decoder_ffn_dim=64, vocab_size=100)
    tmp = tempfile.mkdtemp()
    WhisperEncoder(cfgw).save_pretrained(tmp)
    return tmp


def _compose_whisper_file(model, pcm, batch_size):
    """The reference functions composed as data/loaders.py:177-182 + inference.py:148-211 intend
    (inference.py itself cannot run Whisper models; SURVEY.md finding 5)."""
    from oracle.segma_oracle import file_batches

    t = torch.from_numpy(pcm)
    chunks = []
    for start, n_win, wl in file_batches(len(pcm), 64000, batch_size):
        feats = torch.cat([model.audio_preparation_hook(t[start + i * 63680: start + i * 63680 + wl]) for i in range(n_win)])
        with torch.inference_mode():
            out = model(feats)
        if isinstance(out, dict):
            out = torch.stack([out[f"linear_head_{lab}"] for lab in LABELS], dim=-1)
        if wl < 64000:
            out = out[:, : (wl - 400) // 320 + 1]
        chunks.append(out.reshape(-1, len(LABELS)))
    return torch.cat(chunks)


def models():
    le = MultiLabelEncoder(list(LABELS))
    cs = ConvolutionSettings((320,), (320,), (0,))
    thr = {lab: {"lower_bound": 0.5, "upper_bound": 1.0} for lab in LABELS}
    out = {}
    n, audio_seed, bs = 63680 * 3 + 20000, 11, 2
    pcm = synth.synth_audio(n, audio_seed)
    dims = synth.WHISPER_TEST
    enc_dir = _whisper_dir(dims)
    # surgical_hydra
    cfg = ref_shim.make_config("surgical_hydra", SurgicalHydraConfig(encoder=enc_dir, encoder_layers=[], reduction="weighted",
                                                                      lstm=LSTMConfig(128, 2, True, 0.5), classifier=256))
    m = Models["surgical_hydra"](le, cfg).eval()
    m.load_state_dict(synth.surgical_hydra_state_dict(dims, seed=3), strict=True)
    out["surgical_hydra_logits"] = _compose_whisper_file(m, pcm, bs).numpy()
    # layer subset + average reduction
    cfg2 = ref_shim.make_config("surgical_hydra", SurgicalHydraConfig(encoder=enc_dir, encoder_layers=[2], reduction="average",
                                                                       lstm=LSTMConfig(128, 2, True, 0.5), classifier=256))
    m2 = Models["surgical_hydra"](le, cfg2).eval()
    m2.load_state_dict(synth.surgical_hydra_state_dict(dims, n_mixed_layers=1, seed=8), strict=True)
    out["surgical_hydra_avg_l2_logits"] = _compose_whisper_file(m2, pcm, bs).numpy()
    # hydra_whisper
    cfg3 = ref_shim.make_config("hydra_whisper", HydraWhisperConfig(encoder=enc_dir, lstm=LSTMConfig(128, 2, True, 0.5), classifier=256))
    m3 = Models["hydra_whisper"](le, cfg3).eval()
    m3.load_state_dict(synth.hydra_whisper_state_dict(dims, seed=7), strict=True)
    out["hydra_whisper_logits"] = _compose_whisper_file(m3, pcm, bs).numpy()
    # hubert through the reference's own apply_model_on_audio
    sub = SurgicalHydraLightHuBERTConfig(wav_encoder="none", encoder_layers=[], reduction="weighted", classifier=256, freeze_encoder=True)
    cfg4 = ref_shim.make_config("surgical_hubert_hydra", sub)
    m4 = Models["surgical_hubert_hydra"](le, cfg4, train=False).eval()
    torch.nn.Module.load_state_dict(m4, synth.hubert_hydra_state_dict(synth.HUBERT_BASE, seed=5), strict=True)
    audio = ref_shim.InMemoryAudio()
    audio.add("/mem/a.wav", pcm)
    audio.patch()
    hub = inf.apply_model_on_audio(Path("/mem/a.wav"), m4, cs, "cpu", batch_size=bs)
    out["hubert_logits"] = hub.numpy()
    import torchaudio

    m5 = Models["surgical_hubert_hydra"](le, cfg4, train=False).eval()
    m5.wav2vec2 = torchaudio.models.wavlm_base().eval()
    torch.nn.Module.load_state_dict(m5, synth.hubert_hydra_state_dict(synth.WAVLM_BASE, seed=6), strict=True)
    out["wavlm_logits"] = inf.apply_model_on_audio(Path("/mem/a.wav"), m5, cs, "cpu", batch_size=bs).numpy()
    # thresholds + intervals of the reference on the hubert logits
    for tname, t in (("t50", 0.5), ("t30", 0.3), ("t70", 0.7)):
        th = {lab: {"lower_bound": t, "upper_bound": 1.0} for lab in LABELS}
        mask = inf.apply_thresholds(hub, th, "cpu")
        iv = inf.create_intervals(mask, cs, le)
        out[f"hubert_mask_{tname}"] = mask.numpy()
        out[f"hubert_intervals_{tname}"] = np.array([(LABELS.index(l), s, e) for s, e, l in iv], dtype=np.int64).reshape(-1, 3)
    out["meta"] = np.array([n, audio_seed, bs])
    np.savez(OUT / "models.npz", **out)
    del thr


def intervals():
    """create_intervals edge cases (empty, all-true, single frames, alternating) through the reference."""
    le = MultiLabelEncoder(list(LABELS))
    cs = ConvolutionSettings((320,), (320,), (0,))
    rng = np.random.default_rng(0)
    out = {}
    for i, (nf, p) in enumerate([(1, 1.0), (1, 0.0), (2, 0.5), (7, 0.5), (1024, 0.5), (1025, 0.9), (3000, 0.1), (257, 1.0), (64, 0.0)]):
        mk = rng.random((nf, 4)) < p
        iv = inf.create_intervals(torch.from_numpy(mk), cs, le)
        out[f"mask_{i}"] = mk
        out[f"iv_{i}"] = np.array([(LABELS.index(l), s, e) for s, e, l in iv], dtype=np.int64).reshape(-1, 3)
    alt = np.zeros((50, 4), bool)
    alt[::2, 0] = True
    alt[1::2, 1] = True
    alt[:, 2] = True
    out["mask_alt"] = alt
    iv = inf.create_intervals(torch.from_numpy(alt), cs, le)
    out["iv_alt"] = np.array([(LABELS.index(l), s, e) for s, e, l in iv], dtype=np.int64).reshape(-1, 3)
    np.savez(OUT / "intervals.npz", **out)


def _hubert_reference_model(le, wavlm: bool, dims, seed):
    """The reference's ``SurgicalHydraHubert`` (torchaudio ``wavlm_base`` swapped in for WavLM)."""
    sub = SurgicalHydraLightHuBERTConfig(wav_encoder="none", encoder_layers=[], reduction="weighted", classifier=256, freeze_encoder=True)
    m = Models["surgical_hubert_hydra"](le, ref_shim.make_config("surgical_hubert_hydra", sub), train=False).eval()
    if wavlm:
        import torchaudio

        m.wav2vec2 = torchaudio.models.wavlm_base().eval()
    torch.nn.Module.load_state_dict(m, synth.hubert_hydra_state_dict(dims, seed=seed), strict=True)
    return m


def models_large():
    """HuBERT-base / WavLM-base+ dims on 26 windows + tail (20 800 label decisions each: enough to resolve the
    99.9 % agreement bar) through the reference's own ``apply_model_on_audio``."""
    le = MultiLabelEncoder(list(LABELS))
    cs = ConvolutionSettings((320,), (320,), (0,))
    n, audio_seed, bs = 63680 * 26 + 320 + 8320, 31, 128
    pcm = synth.synth_audio(n, audio_seed)
    audio = ref_shim.InMemoryAudio()
    audio.add("/mem/large.wav", pcm)
    audio.patch()
    out = {"meta": np.array([n, audio_seed, bs])}
    for key, wavlm, dims, seed in (("hubert_logits", False, synth.HUBERT_BASE, 5), ("wavlm_logits", True, synth.WAVLM_BASE, 6)):
        m = _hubert_reference_model(le, wavlm, dims, seed)
        out[key] = inf.apply_model_on_audio(Path("/mem/large.wav"), m, cs, "cpu", batch_size=bs).numpy()
        assert out[key].shape == ((n - 400) // 320 + 1, 4), out[key].shape
    np.savez_compressed(OUT / "models_large.npz", **out)


def whisper_config2():
    """BASELINE config 2's model (Whisper-small dims ``surgical_hydra``, the weights bench.py uses) on one full
    128-window batch + a remainder batch of 8 + the tail, composed from the reference's own classes: the LSTM
    recurrence runs over all 128 windows of the first forward call (surgical_hydra.py:57-60,101)."""
    le = MultiLabelEncoder(list(LABELS))
    dims = synth.WHISPER_SMALL
    enc_dir = _whisper_dir(dims)
    cfg = ref_shim.make_config("surgical_hydra", SurgicalHydraConfig(encoder=enc_dir, encoder_layers=[], reduction="weighted",
                                                                      lstm=LSTMConfig(128, 2, True, 0.5), classifier=256))
    m = Models["surgical_hydra"](le, cfg).eval()
    m.load_state_dict(synth.surgical_hydra_state_dict(dims, seed=0), strict=True)
    n, audio_seed, bs = 63680 * 136 + 33280, 2, 128
    logits = _compose_whisper_file(m, synth.synth_audio(n, audio_seed), bs).numpy()
    assert logits.shape == (136 * 199 + 103, 4), logits.shape
    np.savez_compressed(OUT / "whisper_config2.npz", logits=logits, meta=np.array([n, audio_seed, bs]))


#: the insertion sequences of /root/reference/tests/test_interval.py (every ``Intervals()`` block of its 20 tests),
#: labels as in the tests; the expected lists are NOT copied -- they are produced by running the reference class
INTERVAL_KAT_INPUTS = [
    [],
    [(0, 10, "a")],
    [(0, 10, "a"), (10, 20, "a")],
    [(0, 5, "b"), (5, 10, "b"), (10, 15, "b")],
    [(0, 10, "a"), (5, 15, "a")],
    [(0, 20, "a"), (5, 10, "a")],
    [(0, 10, "a"), (8, 20, "a")],
    [(0, 10, "a"), (15, 25, "a")],
    [(0, 5, "a"), (10, 15, "a"), (20, 25, "a")],
    [(0, 10, "a"), (10, 20, "b")],
    [(0, 15, "a"), (10, 20, "b")],
    [(0, 10, "a"), (5, 15, "b"), (10, 20, "a"), (12, 18, "b")],
    [(0, 10, 1), (10, 20, 1)],
    [(0, 10, 1), (5, 15, 2), (10, 20, 1)],
    [(0, 10, "a"), (5, 15, 1), (10, 20, "a"), (15, 25, 1)],
    [(5, 5, "a"), (5, 5, "a")],
    [(5, 5, "a"), (5, 10, "a")],
    [(20, 30, "a"), (0, 10, "a"), (10, 20, "a")],
    [(15, 20, "b"), (5, 10, "a"), (0, 5, "a"), (10, 15, "b")],
    [(i, i + 10, "a") for i in range(0, 50, 5)],
    [(-10, 0, "a"), (0, 10, "a")],
    [(-20, -10, "a"), (-15, -5, "a")],
    [(0, 1000000, "a"), (1000000, 2000000, "a")],
    [(0, 5, "a"), (10, 15, "a"), (5, 10, "a"), (7, 12, "b")],
    [(0, 20, "a"), (5, 10, "a"), (12, 15, "a")],
    [(0, 10, "a"), (0, 10, "a"), (0, 10, "a")],
    [(0, 10, "a"), (11, 20, "a")],
    [(0, 10, "a-b"), (10, 20, "a-b")],
    [(0, 10, "label with spaces"), (5, 15, "label with spaces")],
    [(0, 10, ""), (10, 20, "")],
]


def postprocess():
    """the reference's ``Intervals`` struct (src/segma/structs/interval.py:8-54) run on the
    insertion sequences of its own tests and on interval lists decoded from random masks (inserted in a shuffled
    order, with nested / overlapping extras); stored as JSON because labels mix ints and strings."""
    import json

    from segma.structs.interval import Intervals

    cs = ConvolutionSettings((320,), (320,), (0,))
    le = MultiLabelEncoder(list(LABELS))
    cases = []
    for seq in INTERVAL_KAT_INPUTS:
        iv = Intervals()
        states = []
        for item in seq:
            iv.add(item)
            states.append([list(t) for t in iv.intervals])
        cases.append({"adds": [list(t) for t in seq], "after_each_add": states, "final": [list(t) for t in iv.intervals]})
    rng = np.random.default_rng(7)
    for nf, p in [(400, 0.5), (2000, 0.85), (3000, 0.2)]:
        mask = rng.random((nf, 4)) < p
        decoded = inf.create_intervals(torch.from_numpy(mask), cs, le)
        extra = [(int(s) + 160, int(e) + 7 * 320, lab) for s, e, lab in decoded[::9]]  # overlapping / bridging rows
        seq = decoded + extra
        order = rng.permutation(len(seq))
        iv = Intervals()
        for j in order:
            iv.add(seq[int(j)])
        cases.append({"adds": [list(seq[int(j)]) for j in order], "final": [list(t) for t in iv.intervals]})
    (OUT / "postprocess.json").write_text(json.dumps({"source": "segma.structs.interval.Intervals", "cases": cases}))


def tuning():
    """``tune_multilabel`` and ``rttm_to_tensor`` of /root/reference/scripts/tune.py (imported with
    ``ruamel`` stubbed) on seeded logits / a seeded RTTM."""
    import importlib.util
    import math
    import types

    if "ruamel" not in sys.modules:
        ru = types.ModuleType("ruamel")
        ru.yaml = types.ModuleType("ruamel.yaml")
        ru.yaml.YAML = object
        sys.modules["ruamel"], sys.modules["ruamel.yaml"] = ru, ru.yaml
    spec = importlib.util.spec_from_file_location("segma_ref_tune", "/root/reference/scripts/tune.py")
    tune = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tune)
    out = {}
    for name, n, seed, precision in (("p10", 50_000, 0, 0.1), ("p100", 50_000, 1, 0.01), ("p10_rare", 3_000, 2, 0.1)):
        g = torch.Generator().manual_seed(seed)
        truth = (torch.rand((n, 4), generator=g) < torch.tensor([0.3, 0.05, 0.5, 0.0])).float()
        logits = (truth * 2 - 1) * 1.5 + torch.randn((n, 4), generator=g) * 1.7
        n_steps = int(1 / precision)
        thresholds = torch.linspace(0, 1, steps=n_steps).round(decimals=int(math.log10(n_steps)))
        tune.n_steps = n_steps  # the function reads the script's module-level n_steps (tune.py:246)
        best = tune.tune_multilabel({"val": {"true": truth, "pred": logits}}, thresholds, list(LABELS))
        out[f"{name}_meta"] = np.array([n, seed, n_steps])
        out[f"{name}_thresholds"] = thresholds.numpy()
        out[f"{name}_best"] = np.array([best[lab]["lower_bound"] for lab in LABELS], dtype=np.float64)
    with tempfile.TemporaryDirectory() as tmp:
        rng = np.random.default_rng(3)
        lines = []
        for _ in range(40):
            s, d = rng.uniform(0, 60), rng.uniform(0.05, 4)
            lab = ["KCHI", "OCH", "MAL", "FEM", "SPEECH"][int(rng.integers(0, 5))]
            lines.append(f"SPEAKER f <NA> {round(float(s), 8)} {round(float(d), 8)} <NA> <NA> {lab} <NA> <NA>")
        p = Path(tmp) / "f.rttm"
        p.write_text("\n".join(lines) + "\n")
        out["rttm_text"] = np.array("\n".join(lines) + "\n")
        out["rttm_tensor"] = tune.rttm_to_tensor(p, list(LABELS)).numpy().astype(np.uint8)
    np.savez_compressed(OUT / "tuning.npz", **out)



if __name__ == "__main__":
    OUT.mkdir(parents=True, exist_ok=True)
    ALL = {"geometry": geometry, "logmel": logmel, "intervals": intervals, "models": models, "models_large": models_large,
           "whisper_config2": whisper_config2, "postprocess": postprocess, "tuning": tuning}
    for name in (sys.argv[1:] or list(ALL)):
        ALL[name]()
    for p in sorted(OUT.glob("*.np*")) + sorted(OUT.glob("*.json")):
        print(p.name, p.stat().st_size)
