"""Deterministic synthetic audio and random-init weights (measurement + parity inputs).

There is no network for datasets or checkpoints, so every measurement and parity run uses
  * audio: ``0.1*N(0,1)`` noise plus sine bursts at 440*k Hz, k = 1..4, following the reference's
    own synthetic-data recipe (/root/reference/scripts/generate_data.py:41-65,89-155), seeded per file;
  * weights: one fp32 ``state_dict`` with the reference's key names, each tensor drawn
    from its own CPU generator seeded by ``crc32(name) ^ seed`` so that the reference model, the oracle
    and the packed fp16 build all load identical values regardless of construction order.
Scales are chosen so activations stay O(1) through the stack and logits spread over a few units,
as in a trained model; they are not the default ``nn.Module`` inits.
"""
from __future__ import annotations

import zlib
from dataclasses import dataclass

import numpy as np
import torch

SAMPLE_RATE = 16_000
DEFAULT_LABELS = ("KCHI", "OCH", "MAL", "FEM")


def synth_audio(n_samples: int, seed: int = 0) -> np.ndarray:
    """Mono float32 PCM in [-1, 1]."""
    rng = np.random.default_rng(seed)
    x = (0.1 * rng.standard_normal(n_samples)).astype(np.float32)
    dur_s = n_samples / SAMPLE_RATE
    n_bursts = max(1, int(dur_s / 2.0))
    starts = rng.uniform(0.0, max(dur_s - 0.05, 0.0), size=n_bursts)
    lens = rng.uniform(0.2, 3.0, size=n_bursts)
    which = rng.integers(1, 5, size=n_bursts)
    for s, d, k in zip(starts, lens, which):
        a = int(s * SAMPLE_RATE)
        b = min(n_samples, a + int(d * SAMPLE_RATE))
        if b <= a:
            continue
        t = np.arange(b - a, dtype=np.float64) / SAMPLE_RATE
        x[a:b] += (0.5 * np.sin(2 * np.pi * 440.0 * k * t)).astype(np.float32)
    return np.clip(x, -1.0, 1.0)


@dataclass(frozen=True)
class WhisperDims:
    d_model: int = 768
    n_layers: int = 12
    ffn: int = 3072
    n_mels: int = 80
    n_ctx: int = 1500

    @property
    def n_heads(self) -> int:
        return self.d_model // 64


WHISPER_SMALL = WhisperDims(768, 12, 3072)
WHISPER_BASE = WhisperDims(512, 6, 2048)
WHISPER_TINY = WhisperDims(384, 4, 1536)
WHISPER_TEST = WhisperDims(128, 2, 256)  # unit-test size


@dataclass(frozen=True)
class LSTMDims:
    hidden_size: int = 128
    num_layers: int = 2
    bidirectional: bool = True


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _normal(name, shape, std, seed, mean=0.0):
    return (torch.randn(shape, generator=_gen(name, seed), dtype=torch.float32) * std + mean).contiguous()


def _linear(sd, name, out_f, in_f, seed, gain=1.0, bias=True):
    sd[name + ".weight"] = _normal(name + ".weight", (out_f, in_f), gain / in_f**0.5, seed)
    if bias:
        sd[name + ".bias"] = _normal(name + ".bias", (out_f,), 0.05, seed)


def _norm(sd, name, dim, seed):
    sd[name + ".weight"] = _normal(name + ".weight", (dim,), 0.1, seed, mean=1.0)
    sd[name + ".bias"] = _normal(name + ".bias", (dim,), 0.05, seed)


def _lstm(sd, prefix, input_size, dims: LSTMDims, seed):
    H = dims.hidden_size
    for layer in range(dims.num_layers):
        in_f = input_size if layer == 0 else H * (2 if dims.bidirectional else 1)
        for suffix in ("", "_reverse") if dims.bidirectional else ("",):
            sd[f"{prefix}weight_ih_l{layer}{suffix}"] = _normal(f"{prefix}weight_ih_l{layer}{suffix}", (4 * H, in_f), 1.0 / in_f**0.5, seed)
            sd[f"{prefix}weight_hh_l{layer}{suffix}"] = _normal(f"{prefix}weight_hh_l{layer}{suffix}", (4 * H, H), 1.0 / H**0.5, seed)
            sd[f"{prefix}bias_ih_l{layer}{suffix}"] = _normal(f"{prefix}bias_ih_l{layer}{suffix}", (4 * H,), 0.1, seed)
            sd[f"{prefix}bias_hh_l{layer}{suffix}"] = _normal(f"{prefix}bias_hh_l{layer}{suffix}", (4 * H,), 0.1, seed)


def _heads(sd, labels, in_f, seed, gain=6.0):
    for lab in labels:
        _linear(sd, f"task_heads.linear_head_{lab}", 1, in_f, seed, gain=gain)


def whisper_encoder_state_dict(dims: WhisperDims, seed: int = 0, prefix: str = "w_encoder.") -> dict:
    sd: dict[str, torch.Tensor] = {}
    d = dims.d_model
    p = prefix
    sd[p + "conv1.weight"] = _normal(p + "conv1.weight", (d, dims.n_mels, 3), 1.5 / (3 * dims.n_mels) ** 0.5, seed)
    sd[p + "conv1.bias"] = _normal(p + "conv1.bias", (d,), 0.05, seed)
    sd[p + "conv2.weight"] = _normal(p + "conv2.weight", (d, d, 3), 1.5 / (3 * d) ** 0.5, seed)
    sd[p + "conv2.bias"] = _normal(p + "conv2.bias", (d,), 0.05, seed)
    sd[p + "embed_positions.weight"] = _normal(p + "embed_positions.weight", (dims.n_ctx, d), 0.3, seed)
    for i in range(dims.n_layers):
        lp = f"{p}layers.{i}."
        _linear(sd, lp + "self_attn.k_proj", d, d, seed, gain=1.5, bias=False)
        _linear(sd, lp + "self_attn.v_proj", d, d, seed)
        _linear(sd, lp + "self_attn.q_proj", d, d, seed, gain=1.5)
        _linear(sd, lp + "self_attn.out_proj", d, d, seed, gain=0.7)
        _norm(sd, lp + "self_attn_layer_norm", d, seed)
        _linear(sd, lp + "fc1", dims.ffn, d, seed)
        _linear(sd, lp + "fc2", d, dims.ffn, seed, gain=0.7)
        _norm(sd, lp + "final_layer_norm", d, seed)
    _norm(sd, p + "layer_norm", d, seed)
    return sd


def surgical_hydra_state_dict(dims: WhisperDims = WHISPER_SMALL, lstm: LSTMDims = LSTMDims(), labels=DEFAULT_LABELS,
                              n_mixed_layers: int | None = None, seed: int = 0) -> dict:
    """Keys of ``SurgicalHydra`` (/root/reference/src/segma/models/whisper/surgical_hydra.py:13-78)."""
    sd = whisper_encoder_state_dict(dims, seed)
    n_mix = dims.n_layers if n_mixed_layers is None else n_mixed_layers
    sd["layer_weights"] = _normal("layer_weights", (n_mix,), 0.5, seed, mean=1.0 / n_mix)
    _lstm(sd, "lstm_shared.", dims.d_model, lstm, seed)
    _heads(sd, labels, lstm.hidden_size * (2 if lstm.bidirectional else 1), seed)
    return sd


def hydra_whisper_state_dict(dims: WhisperDims = WHISPER_TINY, lstm: LSTMDims = LSTMDims(), labels=DEFAULT_LABELS,
                             seed: int = 0) -> dict:
    """Keys of ``HydraWhisper`` (/root/reference/src/segma/models/whisper/hydra.py:20-69)."""
    sd = whisper_encoder_state_dict(dims, seed)
    _lstm(sd, "lstm_shared.", dims.d_model, lstm, seed)
    _heads(sd, labels, lstm.hidden_size * (2 if lstm.bidirectional else 1), seed)
    return sd


@dataclass(frozen=True)
class W2V2Dims:
    d_model: int = 768
    n_layers: int = 12
    ffn: int = 3072
    conv_dim: int = 512
    pos_kernel: int = 128
    pos_groups: int = 16
    wavlm: bool = False
    num_buckets: int = 320

    @property
    def n_heads(self) -> int:
        return self.d_model // 64


HUBERT_BASE = W2V2Dims()
WAVLM_BASE = W2V2Dims(wavlm=True)
W2V2_TEST = W2V2Dims(d_model=128, n_layers=2, ffn=256, conv_dim=128, pos_kernel=16, pos_groups=4)
WAVLM_TEST = W2V2Dims(d_model=128, n_layers=2, ffn=256, conv_dim=128, pos_kernel=16, pos_groups=4, wavlm=True)

W2V2_KERNELS = (10, 3, 3, 3, 3, 2, 2)


def hubert_hydra_state_dict(dims: W2V2Dims = HUBERT_BASE, labels=DEFAULT_LABELS, seed: int = 0) -> dict:
    """Keys of ``SurgicalHydraHubert`` (/root/reference/src/segma/models/hubert/surgical_hydra.py:16-85)
    around torchaudio's wav2vec2 / WavLM module tree."""
    sd: dict[str, torch.Tensor] = {}
    c, d = dims.conv_dim, dims.d_model
    fe = "wav2vec2.feature_extractor."
    for i, k in enumerate(W2V2_KERNELS):
        cin = 1 if i == 0 else c
        # gain ~ sqrt(2)/0.6 keeps the post-GELU scale from collapsing through 7 layers
        sd[f"{fe}conv_layers.{i}.conv.weight"] = _normal(f"{fe}conv_layers.{i}.conv.weight", (c, cin, k), 1.7 / (cin * k) ** 0.5, seed)
    _norm(sd, fe + "conv_layers.0.layer_norm", c, seed)
    enc = "wav2vec2.encoder."
    _norm(sd, enc + "feature_projection.layer_norm", c, seed)
    _linear(sd, enc + "feature_projection.projection", d, c, seed)
    t = enc + "transformer."
    cg = d // dims.pos_groups
    sd[t + "pos_conv_embed.conv.bias"] = _normal(t + "pos_conv_embed.conv.bias", (d,), 0.05, seed)
    v = _normal(t + "pos_conv_embed.conv.v", (d, cg, dims.pos_kernel), 1.0, seed)
    g = _normal(t + "pos_conv_embed.conv.g", (1, 1, dims.pos_kernel), 0.1, seed, mean=(d * cg) ** 0.5 / (cg * dims.pos_kernel) ** 0.5).abs()
    sd[t + "pos_conv_embed.conv.parametrizations.weight.original0"] = g
    sd[t + "pos_conv_embed.conv.parametrizations.weight.original1"] = v
    _norm(sd, t + "layer_norm", d, seed)
    for i in range(dims.n_layers):
        lp = f"{t}layers.{i}."
        if dims.wavlm:
            a = lp + "attention.attention."
            sd[a + "in_proj_weight"] = _normal(a + "in_proj_weight", (3 * d, d), 1.3 / d**0.5, seed)
            sd[a + "in_proj_bias"] = _normal(a + "in_proj_bias", (3 * d,), 0.05, seed)
            _linear(sd, a + "out_proj", d, d, seed, gain=0.7)
            _linear(sd, lp + "attention.gru_rel_pos_linear", 8, 64, seed)
            sd[lp + "attention.gru_rel_pos_const"] = _normal(lp + "attention.gru_rel_pos_const", (1, dims.n_heads, 1, 1), 0.2, seed, mean=1.0)
            if i == 0:
                sd[lp + "attention.rel_attn_embed.weight"] = _normal(lp + "attention.rel_attn_embed.weight", (dims.num_buckets, dims.n_heads), 0.5, seed)
        else:
            for nm, gain in (("q_proj", 1.5), ("k_proj", 1.5), ("v_proj", 1.0), ("out_proj", 0.7)):
                _linear(sd, lp + "attention." + nm, d, d, seed, gain=gain)
        _norm(sd, lp + "layer_norm", d, seed)
        _linear(sd, lp + "feed_forward.intermediate_dense", dims.ffn, d, seed)
        _linear(sd, lp + "feed_forward.output_dense", d, dims.ffn, seed, gain=0.7)
        _norm(sd, lp + "final_layer_norm", d, seed)
    sd["layer_weights"] = torch.full((dims.n_layers,), 1.0 / dims.n_layers)
    _heads(sd, labels, d, seed, gain=1.5)
    return sd
