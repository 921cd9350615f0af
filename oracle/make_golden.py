"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the REFERENCE itself
(/root/reference, imported through oracle/ref_shim.py) on seeded synthetic inputs.

    python oracle/make_golden.py

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


if __name__ == "__main__":
    OUT.mkdir(parents=True, exist_ok=True)
    geometry()
    logmel()
    intervals()
    models()
    for p in sorted(OUT.glob("*.npz")):
        print(p.name, p.stat().st_size)
