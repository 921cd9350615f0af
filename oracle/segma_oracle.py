"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's sliding-window inference path.

Plain torch fp32 (floating-point stages) and numpy / Python integers (geometry, run-length
decode).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module; nothing under ``segma_b200/`` does.

Parity pin: this restatement is checked (tests/test_oracle_vs_reference.py, run where
``/root/reference`` is mounted) against the reference's own classes imported through
``oracle/ref_shim.py`` and against the installed third-party modules that carry the
reference's arithmetic (transformers 5.5.0 ``WhisperFeatureExtractor`` / ``WhisperEncoder``,
torchaudio 2.11.0 wav2vec2 / WavLM components; the reference pins 4.48.3 / 2.9.1 in uv.lock),
and against the committed fixtures in tests/golden/ that were generated from the reference
by ``oracle/make_golden.py``.  The reference's own tests pin only ``ConvolutionSettings``
and ``MultiLabelEncoder``; everything else is pinned by running the
reference code itself.

Every function cites the reference (or third-party) lines it restates; paths are relative
to /root/reference unless prefixed ``site-packages/``.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

SAMPLE_RATE = 16_000
FRAME = 320
N_FFT = 400
HOP = 160
N_MELS = 80
WHISPER_SAMPLES = 480_000
WHISPER_FRAMES = 3000
HEAD_DIM = 64


# ----------------------------------------------------------------------------------------
# geometry (integer)
# ----------------------------------------------------------------------------------------
def rf_start(u: int, kernels, strides, paddings) -> int:
    """src/segma/models/base.py:31-50"""
    jump, off, total = 1, 0, 1
    for k, s, p in zip(kernels, strides, paddings):
        off += p * jump
        jump *= s
    total = jump
    return u * total - off


def rf_end(v: int, kernels, strides, paddings) -> int:
    """src/segma/models/base.py:52-74"""
    jump, back = 1, 0
    for k, s, p in zip(kernels, strides, paddings):
        back += (1 + p - k) * jump
        jump *= s
    return v * jump - back


def rf_size(kernels, strides) -> int:
    """src/segma/models/base.py:76-91"""
    jump, rf = 1, 0
    for k, s in zip(kernels, strides):
        rf += (k - 1) * jump
        jump *= s
    return rf + 1


def file_batches(n_samples: int, win_len: int = 64_000, batch_size: int = 128, step: int | None = None):
    """The forward calls of ``apply_model_on_audio`` (src/segma/inference.py:129-206) as
    ``(start_sample, n_windows, win_len)`` triples: full batches, one remainder batch, the tail.
    ``step`` defaults to the reference's ``win_len - 320``."""
    step = win_len - FRAME if step is None else step
    n_fit = math.floor((n_samples - win_len) / step) + 1 if n_samples >= win_len else 0
    out = []
    n_full = n_fit // batch_size
    for i in range(n_full):
        out.append((i * batch_size * step, batch_size, win_len))
    rem = n_fit - n_full * batch_size
    if rem > 0:
        out.append((n_full * batch_size * step, rem, win_len))
    last = n_fit * step
    if n_samples - last >= 400:
        out.append((last, 1, n_samples - last))
    return out


# ----------------------------------------------------------------------------------------
# Whisper log-mel front end
# ----------------------------------------------------------------------------------------
def hertz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    lin = 3.0 * f / 200.0
    logstep = 27.0 / np.log(6.4)
    return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-300) / 1000.0) * logstep, lin)


def mel_to_hertz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    logstep = np.log(6.4) / 27.0
    return np.where(m >= 15.0, 1000.0 * np.exp(logstep * (m - 15.0)), 200.0 * m / 3.0)


def whisper_mel_filters() -> np.ndarray:
    """(201, 80) float64 slaney-scale, slaney-normalised triangular filters, 0-8 kHz
    (site-packages/transformers/audio_utils.py ``mel_filter_bank`` as called from
    site-packages/transformers/models/whisper/feature_extraction_whisper.py:95-103)."""
    n_bins = 1 + N_FFT // 2
    mel_pts = np.linspace(hertz_to_mel_slaney(0.0), hertz_to_mel_slaney(8000.0), N_MELS + 2)
    edges = mel_to_hertz_slaney(mel_pts)
    fft_freqs = np.linspace(0, SAMPLE_RATE // 2, n_bins)
    diff = np.diff(edges)
    slopes = np.expand_dims(edges, 0) - np.expand_dims(fft_freqs, 1)
    down = -slopes[:, :-2] / diff[:-1]
    up = slopes[:, 2:] / diff[1:]
    fb = np.maximum(np.zeros(1), np.minimum(down, up))
    enorm = 2.0 / (edges[2 : N_MELS + 2] - edges[:N_MELS])
    return fb * np.expand_dims(enorm, 0)


def whisper_logmel(wave: torch.Tensor, mel_filters: np.ndarray | None = None) -> torch.Tensor:
    """``audio_preparation_hook`` of the Whisper-family models
    (src/segma/models/whisper/hydra.py:197-201): pad the 1-D window to 30 s, then
    site-packages/.../feature_extraction_whisper.py:135-164.  Returns ``(80, 3000)`` fp32."""
    mel = whisper_mel_filters() if mel_filters is None else mel_filters
    x = torch.zeros(WHISPER_SAMPLES, dtype=torch.float32)
    n = min(wave.numel(), WHISPER_SAMPLES)
    x[:n] = wave.reshape(-1)[:n].to(torch.float32)
    window = torch.hann_window(N_FFT)
    stft = torch.stft(x, N_FFT, HOP, window=window, return_complex=True)
    power = stft[..., :-1].abs() ** 2
    mel_spec = torch.from_numpy(mel).to(torch.float32).T @ power
    log_spec = torch.clamp(mel_spec, min=1e-10).log10()
    log_spec = torch.maximum(log_spec, log_spec.max() - 8.0)
    return (log_spec + 4.0) / 4.0


# ----------------------------------------------------------------------------------------
# Whisper encoder
# ----------------------------------------------------------------------------------------
def _ln(x, sd, name, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


def _lin(x, sd, name):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def _mha(q, k, v, n_heads, bias=None):
    """softmax(q k^T + bias) v with (B, T, d) inputs; q is already scaled.  Dispatched the way the
    reference's third-party modules do it -- ``F.scaled_dot_product_attention`` with ``scale=1.0``
    (site-packages/transformers/integrations/sdpa_attention.py:92 as selected by modeling_whisper.py:338-350;
    site-packages/torchaudio/models/wav2vec2/components.py:305, wavlm_attention.py:204 with the gated bias as
    ``attn_mask``) -- so the CPU arm of bench.py costs what the reference's CPU path costs instead of
    materialising the (B, H, T, T) score tensor."""
    B, T, d = q.shape
    hd = d // n_heads
    q = q.view(B, T, n_heads, hd).transpose(1, 2)
    k = k.view(B, -1, n_heads, hd).transpose(1, 2)
    v = v.view(B, -1, n_heads, hd).transpose(1, 2)
    o = F.scaled_dot_product_attention(q, k, v, attn_mask=bias, dropout_p=0.0, is_causal=False, scale=1.0)
    return o.transpose(1, 2).reshape(B, T, d)


def whisper_encoder_hidden_states(sd: dict, x: torch.Tensor, prefix: str = "w_encoder.") -> list[torch.Tensor]:
    """``WhisperEncoder.forward(output_hidden_states=True).hidden_states``
    (site-packages/transformers/models/whisper/modeling_whisper.py:593-647; layer 380-412;
    attention 279-355).  Returns ``[embeddings, h_1, ..., h_{N-1}, LN(h_N)]``."""
    p = prefix
    d = sd[p + "conv1.weight"].shape[0]
    n_heads = d // HEAD_DIM
    h = F.gelu(F.conv1d(x, sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1))
    h = F.gelu(F.conv1d(h, sd[p + "conv2.weight"], sd[p + "conv2.bias"], stride=2, padding=1))
    h = h.permute(0, 2, 1) + sd[p + "embed_positions.weight"]
    states = [h]
    n_layers = 0
    while f"{p}layers.{n_layers}.fc1.weight" in sd:
        n_layers += 1
    for i in range(n_layers):
        lp = f"{p}layers.{i}."
        y = _ln(h, sd, lp + "self_attn_layer_norm")
        q = _lin(y, sd, lp + "self_attn.q_proj") * (HEAD_DIM**-0.5)
        k = _lin(y, sd, lp + "self_attn.k_proj")
        v = _lin(y, sd, lp + "self_attn.v_proj")
        h = h + _lin(_mha(q, k, v, n_heads), sd, lp + "self_attn.out_proj")
        y = _ln(h, sd, lp + "final_layer_norm")
        h = h + _lin(F.gelu(_lin(y, sd, lp + "fc1")), sd, lp + "fc2")
        states.append(h)
    states[-1] = _ln(h, sd, p + "layer_norm")
    return states


# ----------------------------------------------------------------------------------------
# LSTM (sequence axis first -- the reference never sets batch_first)
# ----------------------------------------------------------------------------------------
def lstm_seq_first(sd: dict, x: torch.Tensor, prefix: str = "lstm_shared.") -> torch.Tensor:
    """``nn.LSTM(input_size, hidden_size, num_layers, bidirectional)`` in eval mode on an
    input read as ``(seq, batch, feat)`` (src/segma/models/whisper/hydra.py:48-51,81 and
    surgical_hydra.py:57-60,101 -- ``batch_first`` is left False, so ``seq`` is the window
    batch and ``batch`` the frame axis; SURVEY.md finding 6).  Gate order i, f, g, o."""
    layer = 0
    inp = x
    while f"{prefix}weight_ih_l{layer}" in sd:
        outs = []
        for suffix in ("", "_reverse"):
            key = f"{prefix}weight_ih_l{layer}{suffix}"
            if key not in sd:
                continue
            w_ih, w_hh = sd[key], sd[f"{prefix}weight_hh_l{layer}{suffix}"]
            b = sd[f"{prefix}bias_ih_l{layer}{suffix}"] + sd[f"{prefix}bias_hh_l{layer}{suffix}"]
            H = w_hh.shape[1]
            S, N, _ = inp.shape
            h = inp.new_zeros(N, H)
            c = inp.new_zeros(N, H)
            out = inp.new_zeros(S, N, H)
            order = range(S - 1, -1, -1) if suffix else range(S)
            pre = inp @ w_ih.T + b
            for t in order:
                g = pre[t] + h @ w_hh.T
                i_g, f_g, g_g, o_g = g.split(H, dim=-1)
                c = torch.sigmoid(f_g) * c + torch.sigmoid(i_g) * torch.tanh(g_g)
                h = torch.sigmoid(o_g) * torch.tanh(c)
                out[t] = h
            outs.append(out)
        inp = torch.cat(outs, dim=-1)
        layer += 1
    return inp


def _heads(sd: dict, x: torch.Tensor, labels) -> torch.Tensor:
    """``torch.stack([head(x) ...], dim=-1)`` over ``task_heads`` in ``base_labels`` order
    (surgical_hydra.py:107-109, hubert/surgical_hydra.py:101) -> (..., 1, C)."""
    return torch.stack([_lin(x, sd, f"task_heads.linear_head_{lab}") for lab in labels], dim=-1)


def surgical_hydra_forward(sd: dict, x: torch.Tensor, labels, encoder_layers=None, reduction="weighted",
                           n_keep: int = 199) -> torch.Tensor:
    """``SurgicalHydra.forward`` (src/segma/models/whisper/surgical_hydra.py:80-109):
    ``(B, 80, 3000)`` -> ``(B, 199, 1, C)``.  The LSTM runs over the window axis with the
    1500 positions as its batch, so truncating to 199 positions before it is exact."""
    hs = whisper_encoder_hidden_states(sd, x)[1:]
    use = list(range(len(hs))) if not encoder_layers else sorted(i - 1 for i in encoder_layers)
    w = sd["layer_weights"]
    w = torch.softmax(w, dim=0) if reduction == "weighted" else w
    mix = sum(w[j] * hs[i][:, :n_keep] for j, i in enumerate(use))
    out = lstm_seq_first(sd, mix)
    return _heads(sd, out, labels)


def hydra_whisper_forward(sd: dict, x: torch.Tensor, labels, n_keep: int = 199) -> torch.Tensor:
    """``HydraWhisper.forward`` (src/segma/models/whisper/hydra.py:71-87) with the per-head
    dict stacked in ``base_labels`` order -> ``(B, 199, 1, C)``."""
    enc = whisper_encoder_hidden_states(sd, x)[-1][:, :n_keep]
    return _heads(sd, lstm_seq_first(sd, enc), labels)


# ----------------------------------------------------------------------------------------
# wav2vec2 / HuBERT / WavLM
# ----------------------------------------------------------------------------------------
W2V2_KERNELS = (10, 3, 3, 3, 3, 2, 2)
W2V2_STRIDES = (5, 2, 2, 2, 2, 2, 2)


def w2v2_feature_extractor(sd: dict, x: torch.Tensor, prefix: str = "wav2vec2.feature_extractor.",
                           trace: list | None = None) -> torch.Tensor:
    """torchaudio ``FeatureExtractor`` with GroupNorm on layer 0 only
    (site-packages/torchaudio/models/wav2vec2/components.py:77-99,117-143): (B, n) -> (B, T, 512)."""
    h = x.unsqueeze(1)
    for i, s in enumerate(W2V2_STRIDES):
        h = F.conv1d(h, sd[f"{prefix}conv_layers.{i}.conv.weight"], None, stride=s)
        if i == 0:
            c = h.shape[1]
            h = F.group_norm(h, c, sd[f"{prefix}conv_layers.0.layer_norm.weight"],
                             sd[f"{prefix}conv_layers.0.layer_norm.bias"], 1e-5)
        h = F.gelu(h)
        if trace is not None:
            trace.append((f"conv{i}", h.transpose(1, 2)))
    return h.transpose(1, 2)


def _pos_conv_weight(sd: dict, p: str) -> torch.Tensor:
    """weight-norm(dim=2): w = g * v / ||v|| with the norm over dims (0, 1)
    (components.py:194-234; parametrizations.weight.original0 = g, original1 = v)."""
    g = sd[p + "pos_conv_embed.conv.parametrizations.weight.original0"]
    v = sd[p + "pos_conv_embed.conv.parametrizations.weight.original1"]
    return v * (g / v.norm(dim=(0, 1), keepdim=True))


def wavlm_position_bias(rel_attn_embed: torch.Tensor, T: int, num_buckets: int = 320, max_distance: int = 800):
    """Bucketed relative-position bias (site-packages/torchaudio/models/wav2vec2/wavlm_attention.py:85-139):
    returns (n_heads, T, T)."""
    ctx = torch.arange(T)[:, None]
    mem = torch.arange(T)[None, :]
    rel = mem - ctx
    nb = num_buckets // 2
    buckets = (rel > 0).to(torch.long) * nb
    rel = rel.abs()
    max_exact = nb // 2
    is_small = rel < max_exact
    large = max_exact + (
        torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact) * (nb - max_exact)
    ).to(torch.long)
    large = torch.min(large, torch.full_like(large, nb - 1))
    buckets = buckets + torch.where(is_small, rel, large)
    return F.embedding(buckets, rel_attn_embed).permute(2, 0, 1)


def w2v2_encoder_last(sd: dict, feats: torch.Tensor, prefix: str = "wav2vec2.encoder.",
                      trace: list | None = None) -> torch.Tensor:
    """``Encoder.extract_features(x, None)[-1]`` for the base (post-LN) architecture
    (components.py:171-183 projection, 220-234 pos-conv, 421-428 preprocess, 363-401 layer,
    263-310 attention; WavLM attention wavlm_attention.py:166-211).  (B, T, 512) -> (B, T, 768)."""
    p = prefix
    x = _ln(feats, sd, p + "feature_projection.layer_norm")
    x = _lin(x, sd, p + "feature_projection.projection")
    t = p + "transformer."
    d = x.shape[-1]
    n_heads = d // HEAD_DIM
    w = _pos_conv_weight(sd, t)
    k = w.shape[-1]
    pc = F.conv1d(x.transpose(1, 2), w, sd[t + "pos_conv_embed.conv.bias"], padding=k // 2, groups=d // w.shape[1])
    if k % 2 == 0:
        pc = pc[..., :-1]
    if trace is not None:
        trace.append(("proj", x))
    x = x + F.gelu(pc.transpose(1, 2))
    if trace is not None:
        trace.append(("posconv_sum", x))
    x = _ln(x, sd, t + "layer_norm")
    if trace is not None:
        trace.append(("ln0", x))
    wavlm = (t + "layers.0.attention.attention.in_proj_weight") in sd
    pos_bias = None
    if wavlm:
        pos_bias = wavlm_position_bias(sd[t + "layers.0.attention.rel_attn_embed.weight"], x.shape[1])
    i = 0
    while f"{t}layers.{i}.feed_forward.intermediate_dense.weight" in sd:
        lp = f"{t}layers.{i}."
        if wavlm:
            B, T, _ = x.shape
            wi, bi = sd[lp + "attention.attention.in_proj_weight"], sd[lp + "attention.attention.in_proj_bias"]
            qkv = F.linear(x, wi, bi)
            q, kk, v = qkv.split(d, dim=-1)
            xh = x.view(B, T, n_heads, HEAD_DIM).permute(0, 2, 1, 3)
            gl = F.linear(xh, sd[lp + "attention.gru_rel_pos_linear.weight"], sd[lp + "attention.gru_rel_pos_linear.bias"])
            gab = torch.sigmoid(gl.view(B, n_heads, T, 2, 4).sum(-1))
            ga, gb = gab[..., 0:1], gab[..., 1:2]
            gate = ga * (gb * sd[lp + "attention.gru_rel_pos_const"] - 1.0) + 2.0
            bias = gate * pos_bias[None]
            a = _mha(q * (HEAD_DIM**-0.5), kk, v, n_heads, bias=bias)
            a = _lin(a, sd, lp + "attention.attention.out_proj")
        else:
            q = _lin(x, sd, lp + "attention.q_proj") * (HEAD_DIM**-0.5)
            kk = _lin(x, sd, lp + "attention.k_proj")
            v = _lin(x, sd, lp + "attention.v_proj")
            a = _lin(_mha(q, kk, v, n_heads), sd, lp + "attention.out_proj")
        x = _ln(x + a, sd, lp + "layer_norm")
        f = _lin(F.gelu(_lin(x, sd, lp + "feed_forward.intermediate_dense")), sd, lp + "feed_forward.output_dense")
        x = _ln(x + f, sd, lp + "final_layer_norm")
        if trace is not None:
            trace.append((f"layer{i}", x))
        i += 1
    return x


def hubert_hydra_forward(sd: dict, x: torch.Tensor, labels, trace: list | None = None) -> torch.Tensor:
    """``SurgicalHydraHubert.forward`` (src/segma/models/hubert/surgical_hydra.py:87-101):
    ``(B, n_samples)`` -> ``(B, T, 1, C)``; dropout is identity in eval.  ``trace`` collects (stage, tensor)
    pairs for the per-stage diagnosis of the CUDA path (tools/diag_w2v2.py)."""
    feats = w2v2_feature_extractor(sd, x, trace=trace)
    return _heads(sd, w2v2_encoder_last(sd, feats, trace=trace), labels)


# ----------------------------------------------------------------------------------------
# file-level driver, thresholds, intervals
# ----------------------------------------------------------------------------------------
def apply_model_on_audio(pcm: torch.Tensor, forward_windows, n_labels: int, win_len: int = 64_000,
                         batch_size: int = 128, whisper: bool = False) -> torch.Tensor:
    """``apply_model_on_audio`` (src/segma/inference.py:119-211) on in-memory PCM.

    ``forward_windows`` maps ``(B, L)`` waveform windows (``whisper=False``) or
    ``(B, 80, 3000)`` log-mels (``whisper=True``: each window goes through
    ``whisper_logmel`` on its own, as the training loader does at
    src/segma/data/loaders.py:177-182) to ``(B, F, 1, C)`` logits.  For a Whisper-family
    tail of L < win_len samples only the first ``(L-400)//320+1`` frames are kept
    (SURVEY.md A.1 (ii)).  Returns ``(n_frames, C)`` raw logits (concatenation)."""
    pcm = pcm.reshape(-1)
    chunks = []
    with torch.inference_mode():
        for start, n_win, wl in file_batches(pcm.numel(), win_len, batch_size):
            step = win_len - FRAME
            wins = torch.stack([pcm[start + i * step : start + i * step + wl] for i in range(n_win)])
            if whisper:
                feats = torch.stack([whisper_logmel(w) for w in wins])
                out = forward_windows(feats)
                if wl < win_len:
                    out = out[:, : (wl - 400) // FRAME + 1]
            else:
                out = forward_windows(wins)
            chunks.append(out.reshape(-1, n_labels))
    if not chunks:
        return torch.zeros(0, n_labels)
    return torch.cat(chunks, dim=0)


def stitch_mean(window_logits: list[torch.Tensor], frame_offsets: list[int], n_frames: int) -> torch.Tensor:
    """Uniform logit-domain mean over the windows covering each frame (the overlapping-window
    extension of inference.py:209-211; equals concatenation when windows tile the frame grid).
    Accumulates in window order in fp32, then divides by the cover count."""
    C = window_logits[0].shape[-1]
    acc = torch.zeros(n_frames, C, dtype=torch.float32)
    cnt = torch.zeros(n_frames, 1, dtype=torch.float32)
    for w, off in zip(window_logits, frame_offsets):
        acc[off : off + w.shape[0]] += w
        cnt[off : off + w.shape[0]] += 1
    return acc / cnt


def apply_thresholds(logits: torch.Tensor, lower_bounds) -> torch.Tensor:
    """``sigmoid(logit) > lower_bound`` in fp32, strict (src/segma/inference.py:214-234)."""
    return logits.to(torch.float32).sigmoid() > torch.tensor(list(lower_bounds), dtype=torch.float32)


def create_intervals(mask, labels) -> list[tuple[int, int, str]]:
    """Per-label maximal runs of True mapped to ``[320*start, 320*stop)`` samples, label-major
    (src/segma/inference.py:237-263 with ConvolutionSettings((320,),(320,),(0,)), 315-319)."""
    m = np.asarray(mask, dtype=bool)
    out = []
    for c, lab in enumerate(labels):
        col = np.concatenate(([False], m[:, c], [False])) if m.shape[0] else np.zeros(2, bool)
        edges = np.flatnonzero(col[1:] != col[:-1])
        for s, e in zip(edges[0::2], edges[1::2]):
            out.append((max(0, FRAME * int(s)), FRAME * (int(e) - 1) + FRAME - 1 + 1, lab))
    return out


def interval_table(mask, n_labels: int) -> np.ndarray:
    """``create_intervals`` as an int32 ``(n, 3)`` table ``(label_idx, start_sample, end_sample)``."""
    iv = create_intervals(mask, list(range(n_labels)))
    return np.array([(c, s, e) for s, e, c in iv], dtype=np.int64).reshape(-1, 3)


# ----------------------------------------------------------------------------------------
# threshold tuning (scripts/tune.py)
# ----------------------------------------------------------------------------------------
def f1_per_label(y_true: torch.Tensor, y_pred: torch.Tensor) -> np.ndarray:
    """``sklearn.metrics.f1_score(average=None, zero_division=1.0)`` on multilabel indicator matrices:
    2TP / (2TP + FP + FN) per column, 1.0 where the denominator is zero (scripts/tune.py:228-236)."""
    t = np.asarray(y_true) != 0
    p = np.asarray(y_pred) != 0
    tp = (t & p).sum(0).astype(np.float64)
    fp = (~t & p).sum(0).astype(np.float64)
    fn = (t & ~p).sum(0).astype(np.float64)
    den = 2 * tp + fp + fn
    return np.where(den > 0, 2 * tp / np.maximum(den, 1), 1.0)


def tune_multilabel(y_true: torch.Tensor, logits: torch.Tensor, thresholds: torch.Tensor, labels, n_steps: int) -> dict:
    """``tune_multilabel`` (scripts/tune.py:213-256): one F1 pass per grid threshold, best = first maximum."""
    scores = {lab: [] for lab in labels}
    for thresh in thresholds:
        f1 = f1_per_label(y_true, logits.sigmoid() > thresh)
        for i, lab in enumerate(labels):
            scores[lab].append((thresh, f1[i]))
    digits = int(math.log10(n_steps))
    out = {}
    for lab in labels:
        best_t, best_s = None, -1.0
        for t, s in scores[lab]:
            if s > best_s:
                best_t, best_s = t, s
        out[lab] = {"lower_bound": round(float(best_t), digits), "upper_bound": 1.0}
    return out


# ----------------------------------------------------------------------------------------
# extensions beyond the reference: hysteresis, gap merging, minimum duration
# ----------------------------------------------------------------------------------------
def hysteresis_mask(logits: torch.Tensor, offset_cuts, onset_cuts) -> np.ndarray:
    """Sequential definition: on when logit > onset cut, off when logit <= offset cut, else unchanged (off initially)."""
    x = logits.numpy()
    n, C = x.shape
    out = np.zeros((n, C), dtype=bool)
    for c in range(C):
        state = False
        for f in range(n):
            if x[f, c] > onset_cuts[c]:
                state = True
            elif not (x[f, c] > offset_cuts[c]):
                state = False
            out[f, c] = state
    return out


def intervals_reduce(intervals: list) -> list:
    """``Intervals._reduce_per_label`` (src/segma/structs/interval.py:19-45): per label, sort and fuse every
    interval whose start is <= the running end (overlapping, nested or touching); then ``sorted()`` over all labels."""
    by_label: dict = {}
    for s, e, lab in intervals:
        by_label.setdefault(lab, []).append((s, e, lab))
    out = []
    for lab, ivs in by_label.items():
        ivs.sort()
        cur = [ivs[0]]
        for s, e, _ in ivs[1:]:
            if s <= cur[-1][1]:
                cur[-1] = (cur[-1][0], max(cur[-1][1], e), lab)
            else:
                cur.append((s, e, lab))
        out += cur
    return sorted(out)


def postprocess_table(table: np.ndarray, max_gap: int, min_dur: int) -> np.ndarray:
    """Merge rows (file, label, start, end) of the same (file, label) with next.start - end <= max_gap, in order
    (the per-label merge of src/segma/structs/interval.py:19-34 when max_gap = 0), then drop rows shorter than min_dur."""
    out = []
    for row in table.tolist():
        if out and out[-1][0] == row[0] and out[-1][1] == row[1] and row[2] - out[-1][3] <= max_gap:
            out[-1][3] = max(out[-1][3], row[3])
        else:
            out.append(list(row))
    out = [r for r in out if r[3] - r[2] >= min_dur]
    return np.array(out, dtype=np.int64).reshape(-1, 4)
