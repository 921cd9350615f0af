"""Drop-in model classes with the reference's constructor / forward / hook surface, backed by the CUDA engine.

Mirrors ``segma.models.Models`` (/root/reference/src/segma/models/__init__.py:8-15) for the models the
inference entry points accept (names containing "hydra", inference.py:431-432):
  * ``SurgicalHydra``       (models/whisper/surgical_hydra.py:13-109)
  * ``HydraWhisper``        (models/whisper/hydra.py:20-87; its forward returns the per-head dict)
  * ``SurgicalHydraHubert`` (models/hubert/surgical_hydra.py:16-101)
``forward`` takes what the reference takes ((B,80,3000) log-mel or (B,n) waveform) and returns what it
returns; ``audio_preparation_hook`` is the Whisper log-mel (on the GPU) or identity.  Weights come from a
reference ``state_dict`` / Lightning ``.ckpt`` -- there is no training path here.
"""
from __future__ import annotations

from pathlib import Path

import torch

from . import ops
from .config import Config
from .encoders import MultiLabelEncoder
from .engine import WhisperEngine, resolve_device
from .geometry import ConvolutionSettings


class BaseSegmentationModel:
    """Inference-only counterpart of models/base.py:145-169 (no Lightning, no autograd)."""

    family = "base"

    def __init__(self, label_encoder: MultiLabelEncoder, config: Config, weight_loss: bool = False) -> None:
        if not isinstance(label_encoder, MultiLabelEncoder):
            raise ValueError(f"Only MultiLabelEncoder is accepted for {type(self).__name__}.")
        self.label_encoder = label_encoder
        self.config = config
        self.conv_settings = ConvolutionSettings((0,), (0,), (0,))
        self.engine = None
        self.device = torch.device("cuda")
        self.training = False

    # -- torch.nn.Module-like plumbing the reference driver touches (inference.py:438-440) --
    def eval(self):
        self.training = False
        return self

    def to(self, device):
        device = torch.device("cuda" if device == "gpu" else device)
        if device.type != "cuda":
            raise ops.SegmaNativeError("segma_b200 models run on CUDA (sm_100a) only; there is no CPU path")
        self.device = resolve_device(device)
        if self.engine is not None:
            self.engine.to(self.device)  # weights packed by load_from_checkpoint follow the model (inference.py:440)
        return self

    def __call__(self, x):
        return self.forward(x)

    def audio_preparation_hook(self, audio_t):
        return audio_t

    def _require_engine(self):
        if self.engine is None:
            raise RuntimeError(f"{type(self).__name__} has no weights: use from_state_dict() or load_from_checkpoint()")
        return self.engine

    def _build_engine(self, sd: dict):  # pragma: no cover - overridden
        raise NotImplementedError

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        self.engine = self._build_engine({k: v for k, v in state_dict.items()})
        return self

    @classmethod
    def from_state_dict(cls, state_dict, label_encoder, config, **kw):
        return cls(label_encoder, config, **kw).load_state_dict(state_dict)

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, label_encoder, config, train: bool = False, **kw):
        """Lightning ``.ckpt`` (``{"state_dict": ...}``) or a bare ``state_dict`` file (inference.py:435-437)."""
        blob = torch.load(Path(checkpoint_path), map_location="cpu", weights_only=False)
        sd = blob["state_dict"] if isinstance(blob, dict) and "state_dict" in blob else blob
        return cls.from_state_dict(sd, label_encoder, config)


class _WhisperFamily(BaseSegmentationModel):
    family = "whisper"
    kind = ""

    def __init__(self, label_encoder, config, weight_loss: bool = False, loss_f: str = "bce") -> None:
        super().__init__(label_encoder, config, weight_loss)
        self.conv_settings = ConvolutionSettings(kernels=(400, 3, 3), strides=(160, 1, 2), paddings=(200, 1, 1))

    @property
    def n_keep(self) -> int:
        return self.conv_settings.n_windows(self.config.audio.chunk_duration_f, strict=False)

    def _build_engine(self, sd):
        mc = self.config.model.config
        return WhisperEngine(sd, self.label_encoder.base_labels, kind=self.kind,
                             encoder_layers=getattr(mc, "encoder_layers", None),
                             reduction=getattr(mc, "reduction", "weighted"), n_keep=self.n_keep, device=self.device)

    def audio_preparation_hook(self, audio_t):
        """1-D waveform -> (1, 80, 3000) log-mel, the Whisper feature extractor of hydra.py:197-201 on the GPU."""
        x = torch.as_tensor(audio_t, dtype=torch.float32).reshape(-1).to(self.device).contiguous()
        n = min(x.numel(), 480_000 - 400)
        with torch.cuda.device(x.device):
            f32, _ = ops.logmel(x, 1, n, n)
        return f32


class SurgicalHydra(_WhisperFamily):
    kind = "surgical_hydra"

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self._require_engine().forward_features(x)


class HydraWhisper(_WhisperFamily):
    kind = "hydra_whisper"

    def forward(self, x: torch.Tensor) -> dict:
        out = self._require_engine().forward_features(x)  # (B, F, 1, C)
        return {f"linear_head_{lab}": out[..., i] for i, lab in enumerate(self.label_encoder.base_labels)}


class SurgicalHydraHubert(BaseSegmentationModel):
    family = "wav2vec2"

    def __init__(self, label_encoder, config, weight_loss: bool = False, train: bool = True) -> None:
        super().__init__(label_encoder, config, weight_loss)
        self.conv_settings = ConvolutionSettings(
            kernels=(10, 3, 3, 3, 3, 2, 2), strides=(5, 2, 2, 2, 2, 2, 2), paddings=(0, 0, 0, 0, 0, 0, 0)
        )

    def _build_engine(self, sd):
        from .engine_w2v2 import W2V2Engine

        return W2V2Engine(sd, self.label_encoder.base_labels, device=self.device)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self._require_engine().forward_waveforms(x)


Models = {
    "hydra_whisper": HydraWhisper,
    "surgical_hydra": SurgicalHydra,
    "surgical_hubert_hydra": SurgicalHydraHubert,
}

__all__ = ["BaseSegmentationModel", "HydraWhisper", "SurgicalHydra", "SurgicalHydraHubert", "Models"]
