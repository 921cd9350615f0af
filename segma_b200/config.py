"""The reference's configuration contract, read-only for this path.

Field names and YAML layout follow /root/reference/src/segma/config/base.py:10-174 so the
same YAML files load; ``load_config`` (base.py:191-219) is re-done without dacite /
omegaconf: strict recursive dataclass construction plus ``a.b.c=value`` overrides.
The hot path reads ``audio.chunk_duration_s``, ``audio.sample_rate``, ``data.classes``,
``model.name`` and ``model.config.*``.
"""
from __future__ import annotations

import dataclasses
import types
import typing
from dataclasses import asdict, dataclass, field
from pathlib import Path
from typing import Literal

import yaml


@dataclass
class BaseConfig:
    def as_dict(self) -> dict:
        return asdict(self)

    def save(self, file_path) -> None:
        try:
            with Path(file_path).open("w") as f:
                yaml.dump(asdict(self), f, default_flow_style=False, sort_keys=False)
        except IOError as e:
            raise IOError(f"Failed to write configuration to {file_path}: {e}")


@dataclass
class WandbConfig(BaseConfig):
    offline: bool
    project: str
    name: str


@dataclass
class DataConfig(BaseConfig):
    dataset_path: str
    classes: list[str]
    dataset_multiplier: float = 1.0


@dataclass
class AudioConfig(BaseConfig):
    chunk_duration_s: float
    sample_rate: int
    strict_frames: bool

    @property
    def chunk_duration_f(self) -> int:
        return int(self.chunk_duration_s * self.sample_rate)


@dataclass
class DataloaderConfig(BaseConfig):
    num_workers: int


@dataclass
class SchedulerConfig(BaseConfig):
    patience: int


@dataclass
class LSTMConfig(BaseConfig):
    hidden_size: int
    num_layers: int
    bidirectional: int
    dropout: float


@dataclass
class HydraWhisperConfig(BaseConfig):
    encoder: str
    lstm: LSTMConfig
    classifier: int


@dataclass
class SurgicalHydraConfig(BaseConfig):
    encoder: str
    encoder_layers: list[int]
    reduction: Literal["average", "weighted"]
    lstm: LSTMConfig
    classifier: int


@dataclass
class SurgicalHydraLightHuBERTConfig(BaseConfig):
    wav_encoder: str
    encoder_layers: list[int]
    reduction: str
    classifier: int
    freeze_encoder: bool = False


#: sub-config class per model name; only the "hydra" models are accepted by the
#: reference's entry points (inference.py:431-432)
MODEL_CONFIGS = {
    "hydra_whisper": HydraWhisperConfig,
    "surgical_hydra": SurgicalHydraConfig,
    "surgical_hubert_hydra": SurgicalHydraLightHuBERTConfig,
}


@dataclass
class ModelConfig(BaseConfig):
    name: str
    chkp_path: str | None
    config: None | HydraWhisperConfig | SurgicalHydraConfig | SurgicalHydraLightHuBERTConfig = None


@dataclass
class TrainConfig(BaseConfig):
    lr: float
    batch_size: int
    max_epochs: int
    validation_metric: str
    extra_val_metrics: list[str]
    profiler: str | None
    dataloader: DataloaderConfig
    scheduler: SchedulerConfig
    seed: int | None = None


@dataclass
class Config(BaseConfig):
    wandb: WandbConfig
    data: DataConfig
    audio: AudioConfig
    model: ModelConfig
    train: TrainConfig


#: defaults of the model sub-configs shipped as YAML next to the reference's config module
#: (config/surgical_hydra.yml, hydra_whisper.yml, surgical_hubert_hydra.yml)
_LSTM_DEFAULT = {"hidden_size": 128, "num_layers": 2, "bidirectional": True, "dropout": 0.5}
DEFAULT_MODEL_CONFIGS = {
    "surgical_hydra": {
        "encoder": "whisper_base_encoder",
        "encoder_layers": [],
        "reduction": "weighted",
        "lstm": dict(_LSTM_DEFAULT),
        "classifier": 256,
    },
    "hydra_whisper": {"encoder": "whisper_tiny_encoder", "lstm": dict(_LSTM_DEFAULT), "classifier": 256},
    "surgical_hubert_hydra": {
        "wav_encoder": "hubert_base",
        "encoder_layers": [],
        "reduction": "weighted",
        "classifier": 256,
    },
}


def _build(cls, data, where: str):
    """Strict dict -> dataclass (unknown or missing keys raise ``ValueError``)."""
    if not isinstance(data, dict):
        raise ValueError(f"{where}: expected a mapping for {cls.__name__}, got {type(data).__name__}")
    hints = typing.get_type_hints(cls)
    names = {f.name for f in dataclasses.fields(cls)}
    unknown = set(data) - names
    if unknown:
        raise ValueError(f"{where}: unknown field(s) {sorted(unknown)} for {cls.__name__}")
    kwargs = {}
    for f in dataclasses.fields(cls):
        if f.name not in data:
            if f.default is dataclasses.MISSING and f.default_factory is dataclasses.MISSING:
                raise ValueError(f"{where}: missing field '{f.name}' for {cls.__name__}")
            continue
        kwargs[f.name] = _coerce(hints[f.name], data[f.name], f"{where}.{f.name}")
    return cls(**kwargs)


def _coerce(tp, value, where: str):
    origin = typing.get_origin(tp)
    if dataclasses.is_dataclass(tp):
        return _build(tp, value, where)
    if origin in (typing.Union, types.UnionType):
        errors = []
        for alt in typing.get_args(tp):
            try:
                return _coerce(alt, value, where)
            except ValueError as e:
                errors.append(str(e))
        raise ValueError(f"{where}: value {value!r} matches no alternative of {tp}: {errors}")
    if tp is type(None):
        if value is None:
            return None
        raise ValueError(f"{where}: expected null")
    if origin is list:
        if not isinstance(value, (list, tuple)):
            raise ValueError(f"{where}: expected a list")
        (inner,) = typing.get_args(tp)
        return [_coerce(inner, v, f"{where}[{i}]") for i, v in enumerate(value)]
    if origin is Literal:
        if value not in typing.get_args(tp):
            raise ValueError(f"{where}: {value!r} not in {typing.get_args(tp)}")
        return value
    if tp is float and isinstance(value, int) and not isinstance(value, bool):
        return float(value)
    if tp is int and isinstance(value, bool):
        return value  # LSTMConfig.bidirectional is typed int but written as a YAML bool
    if tp in (int, float, str, bool):
        if not isinstance(value, tp) or (tp is int and isinstance(value, bool)):
            raise ValueError(f"{where}: expected {tp.__name__}, got {value!r}")
        return value
    return value


def _apply_override(tree: dict, dotted: str) -> None:
    key, _, raw = dotted.partition("=")
    node = tree
    parts = key.split(".")
    for p in parts[:-1]:
        node = node.setdefault(p, {})
    node[parts[-1]] = yaml.safe_load(raw)


def load_config(config_path, cli_extra_args: list[str] = ()) -> Config:
    """YAML file -> ``Config``.  The model sub-config is taken from the file if present,
    else from ``src/segma/config/<model.name>.yml`` relative to cwd (as the reference does),
    else from the defaults shipped with the reference."""
    with Path(config_path).open("r") as f:
        tree = yaml.safe_load(f)
    model = tree["model"]
    for extra in cli_extra_args:  # a model.name override selects which sub-config file is read
        if extra.split("=", 1)[0] == "model.name":
            _apply_override(tree, extra)
    if "config" not in model:
        side = Path(f"src/segma/config/{model['name']}.yml")
        if side.exists():
            with side.open("r") as f:
                model["config"] = yaml.safe_load(f)
        elif model["name"] in DEFAULT_MODEL_CONFIGS:
            model["config"] = dict(DEFAULT_MODEL_CONFIGS[model["name"]])
        else:
            raise ValueError(f"Model config dict of model {model['name']}, could not be loaded")
    model.setdefault("chkp_path", None)
    for extra in cli_extra_args:
        _apply_override(tree, extra)
    name = tree["model"]["name"]
    cfg = _build(Config, {**tree, "model": {**tree["model"], "config": None}}, "config")
    sub = tree["model"]["config"]
    if sub is not None:
        if name not in MODEL_CONFIGS:
            raise ValueError(f"model '{name}' is not on the multi-label inference path (only 'hydra' models are)")
        cfg.model.config = _build(MODEL_CONFIGS[name], sub, "config.model.config")
    return cfg


def make_config(
    model_name: str,
    model_config: dict | None = None,
    classes=("KCHI", "OCH", "MAL", "FEM"),
    chunk_duration_s: float = 4.0,
) -> Config:
    """Programmatic ``Config`` with the reference's default.yml values (config/default.yml)."""
    sub = dict(DEFAULT_MODEL_CONFIGS[model_name])
    sub.update(model_config or {})
    return Config(
        wandb=WandbConfig(offline=True, project="segma_b200", name="inference"),
        data=DataConfig(dataset_path="none", classes=list(classes)),
        audio=AudioConfig(chunk_duration_s=chunk_duration_s, sample_rate=16_000, strict_frames=False),
        model=ModelConfig(name=model_name, chkp_path=None, config=_build(MODEL_CONFIGS[model_name], sub, "model.config")),
        train=TrainConfig(
            lr=1e-3,
            batch_size=32,
            max_epochs=100,
            validation_metric="loss",
            extra_val_metrics=["loss", "f1_score"],
            profiler=None,
            dataloader=DataloaderConfig(num_workers=8),
            scheduler=SchedulerConfig(patience=3),
        ),
    )
