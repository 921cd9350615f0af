"""TEST INFRASTRUCTURE ONLY -- import shim for the read-only reference checkout.

Makes ``/root/reference/src/segma`` importable in THIS container, where seven of its
third-party imports (lightning, torchmetrics, matplotlib, torchcodec, dacite,
omegaconf, interlap) are not installed.  Only the fixture generators under
``oracle/`` and the ``-m "not gpu"`` cross-checks that are skipped when the
reference is absent may call this; nothing in ``segma_b200/`` imports it, and the
GPU box has no ``/root/reference`` at all.

The stubs carry no arithmetic: ``LightningModule`` is ``torch.nn.Module`` with a
no-op ``save_hyperparameters``/``log``; every other stub is an empty module holding
the imported names.  All numerics come from the reference's own code plus the
installed torch / transformers / torchaudio.
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

REFERENCE_ROOT = Path("/root/reference")


def reference_available() -> bool:
    return (REFERENCE_ROOT / "src" / "segma" / "inference.py").exists()


def _stub(name: str, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def install() -> None:
    """Idempotently install the stubs and put the reference on sys.path."""
    if "segma" in sys.modules and getattr(sys.modules["segma"], "__shimmed__", False):
        return
    if not reference_available():
        raise RuntimeError("reference checkout not present at /root/reference")

    import torch

    # transformers probes find_spec("torchcodec"): import its Whisper classes first.
    from transformers import WhisperFeatureExtractor  # noqa: F401
    from transformers.models.whisper.modeling_whisper import WhisperEncoder  # noqa: F401
    import torchaudio  # noqa: F401

    class LightningModule(torch.nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass

        @classmethod
        def load_from_checkpoint(cls, *a, **k):  # pragma: no cover
            raise NotImplementedError("Lightning is stubbed in the oracle shim")

    def _missing(*a, **k):  # pragma: no cover
        raise NotImplementedError("stubbed third-party symbol")

    if "lightning" not in sys.modules:
        _stub("lightning", LightningModule=LightningModule, LightningDataModule=object)
    if "matplotlib" not in sys.modules:
        mpl = _stub("matplotlib")
        mpl.pyplot = _stub("matplotlib.pyplot")
    if "torchmetrics" not in sys.modules:
        tm = _stub("torchmetrics")
        tm.functional = _stub("torchmetrics.functional")
        tm.functional.classification = _stub(
            "torchmetrics.functional.classification",
            multiclass_auroc=_missing,
            multiclass_f1_score=_missing,
            multiclass_roc=_missing,
            binary_f1_score=_missing,
        )
    if "wandb" not in sys.modules:
        try:
            import wandb  # noqa: F401
        except Exception:
            _stub("wandb")
    if "torchcodec" not in sys.modules:
        tc = _stub("torchcodec")
        tc.decoders = _stub("torchcodec.decoders", AudioDecoder=_missing)
        tc.encoders = _stub("torchcodec.encoders", AudioEncoder=_missing)
    if "dacite" not in sys.modules:
        _stub("dacite", from_dict=_missing, Config=_missing)
    if "omegaconf" not in sys.modules:
        _stub("omegaconf", OmegaConf=types.SimpleNamespace(merge=_missing, from_cli=_missing, to_object=_missing))
    if "interlap" not in sys.modules:
        _stub("interlap", InterLap=_missing)

    src = str(REFERENCE_ROOT / "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    import segma  # noqa: F401

    sys.modules["segma"].__shimmed__ = True


def make_config(model_name: str, sub_config, classes=("KCHI", "OCH", "MAL", "FEM"), chunk_duration_s: float = 4.0):
    """Build the reference's Config dataclasses by hand (load_config needs dacite/omegaconf)."""
    install()
    from segma.config.base import (
        AudioConfig,
        Config,
        DataConfig,
        DataloaderConfig,
        ModelConfig,
        SchedulerConfig,
        TrainConfig,
        WandbConfig,
    )

    return Config(
        wandb=WandbConfig(offline=True, project="oracle", name="oracle"),
        data=DataConfig(dataset_path="none", classes=list(classes)),
        audio=AudioConfig(chunk_duration_s=chunk_duration_s, sample_rate=16_000, strict_frames=False),
        model=ModelConfig(name=model_name, chkp_path=None, config=sub_config),
        train=TrainConfig(
            lr=1e-3,
            batch_size=32,
            max_epochs=1,
            validation_metric="loss",
            extra_val_metrics=["loss"],
            profiler=None,
            dataloader=DataloaderConfig(num_workers=0),
            scheduler=SchedulerConfig(patience=3),
        ),
    )


class InMemoryAudio:
    """Replaces torchcodec on ``segma.inference`` with in-memory PCM (path -> 1-D float32)."""

    def __init__(self):
        self.files: dict[str, "object"] = {}

    def add(self, path, pcm):
        self.files[str(Path(path))] = pcm

    def patch(self):
        install()
        import torch
        import segma.inference as inf
        from segma.utils.io import AudioInfo

        files = self.files

        def get_audio_info(audio_p):
            pcm = files[str(Path(audio_p))]
            return AudioInfo(sample_rate=16_000, n_samples=int(pcm.shape[-1]), n_channels=1)

        def get_samples_in_range(audio_p, start_f, duration_f):
            pcm = files[str(Path(audio_p))]
            end = pcm.shape[-1] if duration_f < 0 else start_f + duration_f
            return torch.as_tensor(pcm[start_f:end]).reshape(1, -1).clone()

        inf.get_audio_info = get_audio_info
        inf.get_samples_in_range = get_samples_in_range
        return inf
