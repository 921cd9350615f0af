"""Host-side contracts kept from the reference: label encoder (tests/test_multi_label_encoder.py of the
reference), RTTM text format (annotation.py:86-104), config loading (config/base.py:191-219), WAV access
(utils/io.py:18-47) and the logit-domain threshold cut."""
import numpy as np
import pytest
import torch
import yaml

from segma_b200.annotation import AudioAnnotation, rttm_line
from segma_b200.config import load_config, make_config
from segma_b200.encoders import MultiLabelEncoder
from segma_b200.io import get_all_samples, get_audio_info, get_samples_in_range, write_wav
from segma_b200.thresholds import logit_cut

LABELS = ("MAL", "FEM", "KCHI", "OCH")


def test_label_encoder_maps_and_inverse():
    le = MultiLabelEncoder(LABELS)
    assert le.labels == LABELS and le.base_labels == LABELS and len(le) == 4 and le.n_labels == 4
    for i, lab in enumerate(LABELS):
        assert le.transform(lab) == i == le(lab)
        assert le.inv_transform(i) == lab
    with pytest.raises(ValueError):
        le.inv_transform(4)
    with pytest.raises(ValueError):
        le.inv_transform(-1)
    with pytest.raises(KeyError):
        le.transform("XXX")


def test_label_encoder_one_hot_and_contains():
    le = MultiLabelEncoder(LABELS)
    assert le.one_hot("FEM").tolist() == [0, 1, 0, 0]
    assert le.one_hot(("MAL", "OCH")).tolist() == [1, 0, 0, 1]
    assert le.one_hot(()).tolist() == [0, 0, 0, 0]
    assert le.i_to_one_hot(2).tolist() == [0, 0, 1, 0]
    assert "KCHI" in le and "nope" not in le
    with pytest.raises(ValueError):
        ("MAL", "FEM") in le  # noqa: B015


def test_rttm_format():
    assert rttm_line("file_a", 320, 960, "KCHI") == "SPEAKER file_a <NA> 0.02 0.04 <NA> <NA> KCHI <NA> <NA>"
    assert rttm_line("f", 0, 57_599_680, "FEM") == "SPEAKER f <NA> 0.0 3599.98 <NA> <NA> FEM <NA> <NA>"
    a = AudioAnnotation.from_rttm(rttm_line("u", 63680, 64000 + 12345, "OCH"))
    assert (a.uid, a.label) == ("u", "OCH") and a.start_time_s == 3.98 and a.duration_s == round(12665 / 16000, 8)


def test_config_default_yaml(tmp_path):
    cfg_d = {
        "wandb": {"offline": False, "project": "p", "name": "n"},
        "data": {"dataset_path": "d", "classes": ["KCHI", "OCH", "MAL", "FEM"]},
        "audio": {"chunk_duration_s": 4.0, "sample_rate": 16000, "strict_frames": False},
        "model": {"name": "surgical_hydra"},
        "train": {"lr": 0.001, "batch_size": 32, "max_epochs": 100, "validation_metric": "loss",
                  "extra_val_metrics": ["loss", "f1_score"], "profiler": None, "dataloader": {"num_workers": 8},
                  "scheduler": {"patience": 3}},
    }
    p = tmp_path / "c.yml"
    p.write_text(yaml.safe_dump(cfg_d))
    cfg = load_config(p)
    assert cfg.audio.chunk_duration_f == 64000 and cfg.model.name == "surgical_hydra"
    assert cfg.model.config.lstm.hidden_size == 128 and cfg.model.config.reduction == "weighted"
    cfg2 = load_config(p, ["model.name=hydra_whisper", "audio.chunk_duration_s=2.0"])
    assert cfg2.model.name == "hydra_whisper" and cfg2.audio.chunk_duration_f == 32000
    # round trip through save()
    cfg.save(tmp_path / "out.yml")
    again = load_config(tmp_path / "out.yml")
    assert again.as_dict() == cfg.as_dict()
    # strictness: unknown key is an error
    cfg_d["audio"]["bogus"] = 1
    p.write_text(yaml.safe_dump(cfg_d))
    with pytest.raises(ValueError):
        load_config(p)


def test_make_config_models():
    for name in ("surgical_hydra", "hydra_whisper", "surgical_hubert_hydra"):
        cfg = make_config(name)
        assert cfg.model.name == name and cfg.data.classes == ["KCHI", "OCH", "MAL", "FEM"]


@pytest.mark.parametrize("subtype,tol", [("float32", 0.0), ("int16", 1.0 / 32768)])
def test_wav_io_roundtrip(tmp_path, subtype, tol):
    rng = np.random.default_rng(0)
    x = (0.5 * rng.standard_normal(48_000)).clip(-1, 1).astype(np.float32)
    p = tmp_path / "a.wav"
    write_wav(p, x, subtype=subtype)
    info = get_audio_info(p)
    assert (info.sample_rate, info.n_samples, info.n_channels) == (16000, 48000, 1)
    full = get_all_samples(p)
    assert full.shape == (1, 48000) and full.dtype == torch.float32
    assert np.abs(full.numpy()[0] - x).max() <= tol
    part = get_samples_in_range(p, 1000, 5000)
    assert torch.equal(part, full[:, 1000:6000])
    rest = get_samples_in_range(p, 40_000, -1)
    assert torch.equal(rest, full[:, 40_000:])


def test_wav_io_zero_file(tmp_path):
    """The reference's io fixture: 3 minutes of float32 zeros at 16 kHz (tests/test_io.py:10-24)."""
    p = tmp_path / "00.wav"
    write_wav(p, np.zeros(16000 * 180, dtype=np.float32))
    assert get_audio_info(p).n_samples == 2_880_000
    assert get_samples_in_range(p, 0, 16000).abs().max() == 0


def test_in_memory_audio():
    x = np.arange(10, dtype=np.float32)
    assert get_audio_info(x).n_samples == 10
    assert get_samples_in_range(x, 2, 3).tolist() == [[2.0, 3.0, 4.0]]


@pytest.mark.parametrize("t", [0.5, 0.3, 0.7, 0.05, 0.95, 0.999])
def test_logit_cut_is_exact(t):
    cut = torch.tensor(logit_cut(t))
    ulps = torch.tensor([cut, torch.nextafter(cut, torch.tensor(float("inf"))), torch.nextafter(cut, torch.tensor(float("-inf")))])
    want = ulps.sigmoid() > torch.tensor(t)
    assert want.tolist() == [False, True, False]
    x = torch.linspace(-12, 12, 200_001)
    assert torch.equal(x.sigmoid() > torch.tensor(t), x > cut)


def test_logit_cut_degenerate():
    assert logit_cut(1.0) == float("inf") and logit_cut(-0.1) == float("-inf")
    # sigmoid(x) > 0.5 is not x > 0
    assert logit_cut(0.5) > 0.0


# ---- error behaviour of the driver entry point (inference.py:398-432), checked before any CUDA work ------------
def _write_cfg(tmp_path, model_name="surgical_hubert_hydra"):
    cfg = {
        "wandb": {"offline": True, "project": "p", "name": "n"},
        "data": {"dataset_path": "d", "classes": ["KCHI", "OCH", "MAL", "FEM"]},
        "audio": {"chunk_duration_s": 4.0, "sample_rate": 16000, "strict_frames": False},
        "model": {"name": model_name},
        "train": {"lr": 0.001, "batch_size": 32, "max_epochs": 1, "validation_metric": "loss", "extra_val_metrics": [],
                  "profiler": None, "dataloader": {"num_workers": 0}, "scheduler": {"patience": 3}},
    }
    p = tmp_path / "cfg.yml"
    p.write_text(yaml.safe_dump(cfg))
    return p


def test_driver_errors_match_the_reference(tmp_path):
    from segma_b200.inference import get_list_of_files_to_process, run_inference_on_audios

    cfg = _write_cfg(tmp_path)
    wavs = tmp_path / "wav"
    wavs.mkdir()
    ckpt = tmp_path / "best.ckpt"
    ckpt.write_bytes(b"")
    kw = dict(config=cfg, uris=None, output=tmp_path / "out", batch_size=2)
    with pytest.raises(ValueError):  # missing checkpoint
        run_inference_on_audios(wavs=wavs, checkpoint=tmp_path / "nope.ckpt", thresholds=None, **kw)
    with pytest.raises(ValueError):  # thresholds path that does not exist
        run_inference_on_audios(wavs=wavs, checkpoint=ckpt, thresholds=tmp_path / "nope.yml", **kw)
    with pytest.raises(FileNotFoundError):  # wavs folder
        run_inference_on_audios(wavs=tmp_path / "nowav", checkpoint=ckpt, thresholds=None, **kw)
    bad = _write_cfg(tmp_path, "whisperidou")
    with pytest.raises(ValueError):  # only the multi-label ("hydra") models are accepted
        run_inference_on_audios(config=bad, uris=None, wavs=wavs, checkpoint=ckpt, output=tmp_path / "o", thresholds=None, batch_size=2)
    # file listing: sorted glob, or the uris file
    for name in ("b.wav", "a.wav", "c.txt"):
        (wavs / name).write_bytes(b"")
    files, n = get_list_of_files_to_process(wavs)
    assert [f.name for f in files] == ["a.wav", "b.wav"] and n == 2
    (tmp_path / "uris.txt").write_text("b\nzz\n")
    files, n = get_list_of_files_to_process(wavs, uris=tmp_path / "uris.txt")
    assert [f.name for f in files] == ["b.wav", "zz.wav"] and n == 2


def test_wavlm_relative_bias_is_the_toeplitz_form_of_the_table():
    """The (H, 2T-1) vector handed to segma_attention_rel reproduces every entry of the (H, T, T) table."""
    import torch

    from segma_b200.engine_w2v2 import wavlm_position_bias, wavlm_relative_bias

    emb = torch.randn(320, 12, generator=torch.Generator().manual_seed(3))
    for T in (7, 199, 349):
        pb = wavlm_position_bias(emb, T, 320)
        rel = wavlm_relative_bias(emb, T, 320)
        assert rel.shape == (12, 2 * T - 1)
        idx = torch.arange(T)[None, :] - torch.arange(T)[:, None] + T - 1
        assert torch.equal(rel[:, idx], pb)


def test_nan_threshold_activates_nothing():
    """``sigmoid(x) > nan`` is False for every frame in the reference (inference.py:228-234)."""
    import math

    from segma_b200.thresholds import logit_cut

    assert logit_cut(float("nan")) == math.inf
    assert logit_cut(-0.1) == -math.inf
    assert logit_cut(1.0) == math.inf


def test_positional_conv_weight_from_any_checkpoint_flavour():
    """weight-norm parametrisation keys of current torch, ``weight_g`` / ``weight_v`` of older checkpoints, or a plain
    ``weight``: all give the same effective kernel (torchaudio components.py:194-234)."""
    import torch

    from segma_b200.engine_w2v2 import _pos_conv_weight

    g = torch.rand(1, 1, 8) + 0.5
    v = torch.randn(16, 4, 8)
    w = v * (g / v.norm(dim=(0, 1), keepdim=True))
    p = "enc.pos_conv_embed.conv."
    a = _pos_conv_weight({p + "parametrizations.weight.original0": g, p + "parametrizations.weight.original1": v}, p)
    b = _pos_conv_weight({p + "weight_g": g, p + "weight_v": v}, p)
    c = _pos_conv_weight({p + "weight": w}, p)
    assert torch.equal(a, w) and torch.equal(b, w) and torch.equal(c, w)
    import pytest

    with pytest.raises(KeyError):
        _pos_conv_weight({}, p)
