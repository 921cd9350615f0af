"""Device-side execution of ``SurgicalHydraHubert.forward`` for wav2vec2 / HuBERT / WavLM encoders.

Follows /root/reference/src/segma/models/hubert/surgical_hydra.py:87-101 through torchaudio's module tree
(site-packages/torchaudio/models/wav2vec2/components.py: FeatureExtractor 117-143, FeatureProjection 171-183,
ConvolutionalPositionalEmbedding 194-234, Transformer._preprocess 421-428, EncoderLayer 363-401, SelfAttention
263-310; wavlm_attention.py 166-211), fused with the windowing of inference.py:148-152: raw PCM on the
device in, frame logits on the file timeline out.  All arithmetic is libsegma_b200 kernels:
layer 0 (conv + GroupNorm + GELU) in one kernel pair, layers 1-6 and the grouped positional convolution as
implicit GEMMs on the tcgen05 kernel, post-LN transformer layers, per-label heads.
"""
from __future__ import annotations

import math

import torch

from . import ops
from .engine import move_to, resolve_device

HEAD_DIM = 64
CONV_KERNELS = (10, 3, 3, 3, 3, 2, 2)
CONV_STRIDES = (5, 2, 2, 2, 2, 2, 2)


def _f16(t, device):
    return t.detach().to(device=device, dtype=torch.float16).contiguous()


def _f32(t, device):
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def conv_lengths(n_samples: int) -> list[int]:
    out, t = [], n_samples
    for k, s in zip(CONV_KERNELS, CONV_STRIDES):
        t = (t - k) // s + 1 if t >= k else 0
        out.append(t)
    return out


def _pos_conv_weight(sd: dict, p: str) -> torch.Tensor:
    """The positional convolution's effective kernel.  torchaudio wraps it in weight-norm over dim 2
    (components.py:194-234): checkpoints hold ``parametrizations.weight.original0/1`` (g, v) with current torch,
    ``weight_g`` / ``weight_v`` when saved by older torch / torchaudio (torch's load hook converts those), or a plain
    ``weight`` after ``remove_weight_norm`` / script export.  w = g * v / ||v||, the norm over dims (0, 1)."""
    for kg, kv in (("parametrizations.weight.original0", "parametrizations.weight.original1"), ("weight_g", "weight_v")):
        if p + kg in sd and p + kv in sd:
            g, v = sd[p + kg].float(), sd[p + kv].float()
            return v * (g / v.norm(dim=(0, 1), keepdim=True))
    if p + "weight" in sd:
        return sd[p + "weight"].float()
    raise KeyError(f"no positional-convolution weight under '{p}' (parametrizations.weight.original0/1, weight_g/v or weight)")


def _even(v: int) -> int:
    return v + (v & 1)


def wavlm_position_bias(rel_attn_embed: torch.Tensor, T: int, num_buckets: int, max_distance: int = 800) -> torch.Tensor:
    """(n_heads, T, T) bucketed relative-position bias table (wavlm_attention.py:85-139); a weight-derived
    constant, built once per sequence length on the host."""
    ctx = torch.arange(T)[:, None]
    mem = torch.arange(T)[None, :]
    rel = mem - ctx
    nb = num_buckets // 2
    buckets = (rel > 0).to(torch.long) * nb
    rel = rel.abs()
    max_exact = nb // 2
    large = max_exact + (torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact) * (nb - max_exact)).to(torch.long)
    large = torch.min(large, torch.full_like(large, nb - 1))
    buckets = buckets + torch.where(rel < max_exact, rel, large)
    return torch.nn.functional.embedding(buckets, rel_attn_embed.detach().float().cpu()).permute(2, 0, 1).contiguous()


def wavlm_relative_bias(rel_attn_embed: torch.Tensor, T: int, num_buckets: int, max_distance: int = 800) -> torch.Tensor:
    """(n_heads, 2T-1) Toeplitz form of ``wavlm_position_bias``: entry ``j - i + T - 1`` is the bias of key j for
    query i (the bucket depends on j - i only, wavlm_attention.py:107-139)."""
    pb = wavlm_position_bias(rel_attn_embed, T, num_buckets, max_distance)
    return torch.cat([pb[:, 1:, 0].flip(1), pb[:, 0, :]], dim=1).contiguous()  # offsets -(T-1)..-1, then 0..T-1


class W2V2Engine:
    def __init__(self, sd: dict, labels, device="cuda", prefix: str = "wav2vec2."):
        dev = self.device = resolve_device(device)
        with torch.cuda.device(dev):
            ops.device_check()
        self.labels = tuple(labels)
        fe = prefix + "feature_extractor."
        w0 = sd[fe + "conv_layers.0.conv.weight"]
        self.C = C = w0.shape[0]
        assert C % 128 == 0, "conv feature dimension must be a multiple of 128"
        self.conv0_w = _f32(w0.reshape(C, CONV_KERNELS[0]), dev)
        self.gn_g = _f32(sd[fe + "conv_layers.0.layer_norm.weight"], dev)
        self.gn_b = _f32(sd[fe + "conv_layers.0.layer_norm.bias"], dev)
        self.conv_w = [_f16(sd[f"{fe}conv_layers.{i}.conv.weight"].permute(0, 2, 1).reshape(C, -1), dev) for i in range(1, 7)]
        enc = prefix + "encoder."
        self.proj_ln_g = _f32(sd[enc + "feature_projection.layer_norm.weight"], dev)
        self.proj_ln_b = _f32(sd[enc + "feature_projection.layer_norm.bias"], dev)
        self.proj_w = _f16(sd[enc + "feature_projection.projection.weight"], dev)
        self.proj_b = _f32(sd[enc + "feature_projection.projection.bias"], dev)
        self.d = d = self.proj_w.shape[0]
        self.n_heads = d // HEAD_DIM
        assert d % 128 == 0
        t = enc + "transformer."
        # positional convolution: fold weight-norm, then lay the grouped kernel out for N tiles of whole groups
        w = _pos_conv_weight(sd, t + "pos_conv_embed.conv.")  # (d, cg, K)
        cg, K = w.shape[1], w.shape[2]
        self.pos_k = K
        self.pos_cg = cg
        self.pos_pad = K // 2
        bn = next((b for b in (192, 256, 128) if d % b == 0 and b % cg == 0), None)
        assert bn is not None, f"no N tile fits positional-conv groups of {cg} channels in d={d}"
        self.pos_bn = bn
        wg = torch.zeros((d, K, bn), dtype=torch.float32)
        for co in range(d):
            gl = (co % bn) // cg  # group index inside the N tile
            wg[co, :, gl * cg:(gl + 1) * cg] = w[co].T
        self.pos_w = _f16(wg.reshape(d, K * bn), dev)
        self.pos_b = _f32(sd[t + "pos_conv_embed.conv.bias"], dev)
        self.ln0_g = _f32(sd[t + "layer_norm.weight"], dev)
        self.ln0_b = _f32(sd[t + "layer_norm.bias"], dev)
        self.wavlm = (t + "layers.0.attention.attention.in_proj_weight") in sd
        scale = HEAD_DIM**-0.5
        self.layers = []
        i = 0
        while f"{t}layers.{i}.feed_forward.intermediate_dense.weight" in sd:
            lp = f"{t}layers.{i}."
            if self.wavlm:
                wi, bi = sd[lp + "attention.attention.in_proj_weight"].clone().float(), sd[lp + "attention.attention.in_proj_bias"].clone().float()
                wi[:d] *= scale
                bi[:d] *= scale
                wo, bo = sd[lp + "attention.attention.out_proj.weight"], sd[lp + "attention.attention.out_proj.bias"]
                extra = dict(gate_w=_f32(sd[lp + "attention.gru_rel_pos_linear.weight"], dev),
                             gate_b=_f32(sd[lp + "attention.gru_rel_pos_linear.bias"], dev),
                             gate_c=_f32(sd[lp + "attention.gru_rel_pos_const"].reshape(-1), dev))
            else:
                wi = torch.cat([sd[lp + "attention.q_proj.weight"] * scale, sd[lp + "attention.k_proj.weight"], sd[lp + "attention.v_proj.weight"]])
                bi = torch.cat([sd[lp + "attention.q_proj.bias"] * scale, sd[lp + "attention.k_proj.bias"], sd[lp + "attention.v_proj.bias"]])
                wo, bo = sd[lp + "attention.out_proj.weight"], sd[lp + "attention.out_proj.bias"]
                extra = {}
            self.layers.append(dict(
                wqkv=_f16(wi, dev), bqkv=_f32(bi, dev), wo=_f16(wo, dev), bo=_f32(bo, dev),
                ln1_g=_f32(sd[lp + "layer_norm.weight"], dev), ln1_b=_f32(sd[lp + "layer_norm.bias"], dev),
                w1=_f16(sd[lp + "feed_forward.intermediate_dense.weight"], dev), b1=_f32(sd[lp + "feed_forward.intermediate_dense.bias"], dev),
                w2=_f16(sd[lp + "feed_forward.output_dense.weight"], dev), b2=_f32(sd[lp + "feed_forward.output_dense.bias"], dev),
                ln2_g=_f32(sd[lp + "final_layer_norm.weight"], dev), ln2_b=_f32(sd[lp + "final_layer_norm.bias"], dev), **extra))
            i += 1
        self.ffn = self.layers[0]["w1"].shape[0]
        if self.wavlm:
            self.rel_embed = sd[t + "layers.0.attention.rel_attn_embed.weight"]
            self._pos_bias: dict[int, torch.Tensor] = {}
        self.head_w = _f32(torch.cat([sd[f"task_heads.linear_head_{lab}.weight"] for lab in labels], dim=0), dev)
        self.head_b = _f32(torch.cat([sd[f"task_heads.linear_head_{lab}.bias"] for lab in labels], dim=0), dev)
        self._ws: dict[tuple[int, int], dict] = {}
        #: diagnostics only (tools/diag_w2v2.py): when a list, every stage appends (name, copy of its output)
        self.trace: list | None = None

    def _tr(self, name: str, t: torch.Tensor) -> None:
        if self.trace is not None:
            self.trace.append((name, t.detach().float().clone()))

    def _pos_bias_for(self, T: int) -> torch.Tensor:
        if T not in self._pos_bias:
            pb = wavlm_position_bias(self.rel_embed, T, self.rel_embed.shape[0])
            ld = (T + 3) // 4 * 4  # rows padded to 16 bytes for the tcgen05 kernel's 128-bit loads
            padded = torch.zeros((pb.shape[0], T, ld), dtype=torch.float32, device=self.device)
            padded[:, :, :T] = pb.to(self.device)
            torch.cuda.current_stream(self.device).synchronize()  # cached for every stream that runs this engine later
            self._pos_bias[T] = padded[:, :, :T]  # (H, T, T) view with row stride ld
        return self._pos_bias[T]

    REL_BIAS_MAX_T = 1024  # segma_attention_rel keeps T + 128 floats of the vector in shared memory

    def _rel_bias_for(self, T: int) -> torch.Tensor:
        key = -T
        if key not in self._pos_bias:
            self._pos_bias[key] = wavlm_relative_bias(self.rel_embed, T, self.rel_embed.shape[0]).to(self.device)
            torch.cuda.current_stream(self.device).synchronize()  # cached for every stream that runs this engine later
        return self._pos_bias[key]

    def _workspace(self, n: int, win_len: int, slot: int = 0) -> dict:
        key = (n, win_len, slot)
        if key in self._ws:
            return self._ws[key]
        if len(self._ws) > 8:  # (full batch, remainder, tail) x 2 slots (+ slack); drop the rest
            self._ws.clear()
        dev, C, d = self.device, self.C, self.d
        lens = conv_lengths(win_len)
        T = lens[-1]
        ws = {"lens": lens, "T": T}
        ws["ss"] = torch.empty((n, C, 2), dtype=torch.float32, device=dev)
        ws["act"] = [torch.zeros((n, _even(t), C), dtype=torch.float16, device=dev) for t in lens[:-1]]
        ws["feat"] = torch.empty((n * T, C), dtype=torch.float32, device=dev)
        ws["feat_f16"] = torch.empty((n * T, C), dtype=torch.float16, device=dev)
        rows_p = (T + self.pos_k + 7) // 8 * 8
        ws["xp"] = torch.zeros((n, rows_p, d), dtype=torch.float16, device=dev)
        ws["x0"] = torch.empty((n * T, d), dtype=torch.float32, device=dev)
        ws["tmp"] = torch.empty((n * T, d), dtype=torch.float32, device=dev)
        ws["x"] = torch.empty((n * T, d), dtype=torch.float32, device=dev)
        ws["x_f16"] = torch.empty((n * T, d), dtype=torch.float16, device=dev)
        ws["qkv"] = torch.empty((n * T, 3 * d), dtype=torch.float16, device=dev)
        ws["att"] = torch.empty((n * T, d), dtype=torch.float16, device=dev)
        ws["h1"] = torch.empty((n * T, self.ffn), dtype=torch.float16, device=dev)
        if self.wavlm:
            ws["gate"] = torch.empty((n, self.n_heads, T), dtype=torch.float32, device=dev)
        self._ws[key] = ws
        return ws

    def to(self, device) -> "W2V2Engine":
        """Move the packed weights to another CUDA device; workspaces and bias tables are re-created there on demand."""
        device = resolve_device(device)
        if device != self.device:
            for k, v in vars(self).items():
                if k not in ("_ws", "_pos_bias", "device", "rel_embed"):
                    setattr(self, k, move_to(v, device))
            self._ws, self.device = {}, device
            if self.wavlm:
                self._pos_bias = {}
            with torch.cuda.device(device):
                ops.device_check()
        return self

    def forward_pcm(self, pcm: torch.Tensor, start: int, n: int, win_len: int, step: int, logits: torch.Tensor,
                    frame_offset: int, step_frames: int, n_keep: int | None = None, slot: int = 0) -> None:
        """Windows ``pcm[start + i*step : +win_len]``, i < n -> ``logits[frame_offset + i*step_frames + r]``, r < n_keep."""
        with torch.cuda.device(self.device):
            self._forward_pcm(pcm, start, n, win_len, step, logits, frame_offset, step_frames, n_keep, slot)

    def forward_windows(self, pcm: torch.Tensor, win_offsets: torch.Tensor, n: int, win_len: int, logits: torch.Tensor,
                        frame_offsets: torch.Tensor, n_keep: int | None = None, slot: int = 0) -> None:
        """Windows ``pcm[win_offsets[i] : +win_len]`` -> ``logits[frame_offsets[i] + r]``, r < n_keep, for int64 device
        tables of n offsets: the windows of this model family are independent of each other (no LSTM over the window
        axis), so windows of several files can share one forward call."""
        with torch.cuda.device(self.device):
            self._forward_pcm(pcm, 0, n, win_len, 0, logits, 0, 0, n_keep, slot, win_offsets, frame_offsets)

    def _forward_pcm(self, pcm, start, n, win_len, step, logits, frame_offset, step_frames, n_keep, slot,
                     win_offsets=None, frame_offsets=None) -> None:
        ws = self._workspace(n, win_len, slot)
        lens, T, C, d = ws["lens"], ws["T"], self.C, self.d
        if T <= 0:
            return
        view = pcm[start:]
        act = ws["act"]
        ops.w2v2_layer0(view, n, win_len, step, self.conv0_w, self.gn_g, self.gn_b, ws["ss"], act[0], win_offsets=win_offsets)
        self._tr("conv0", act[0][:, : lens[0]])
        for i in range(1, 7):
            src = act[i - 1]
            if i < 6:
                ops.conv1d_tm(src, self.conv_w[i - 1], None, CONV_KERNELS[i], CONV_STRIDES[i], lens[i], gelu=True,
                              out=act[i], out_batch_rows=act[i].shape[1])
                self._tr(f"conv{i}", act[i][:, : lens[i]])
            else:
                ops.conv1d_tm(src, self.conv_w[i - 1], None, CONV_KERNELS[i], CONV_STRIDES[i], T, gelu=True,
                              out=ws["feat"].view(n, T, C))
        # feature projection, positional convolution, pre-encoder LayerNorm
        ops.layernorm(ws["feat"], self.proj_ln_g, self.proj_ln_b, out_f16=ws["feat_f16"])
        ops.linear(ws["feat_f16"], self.proj_w, self.proj_b, out=ws["x0"])
        xp = ws["xp"]
        ops.gemm_raw(ws["feat_f16"].data_ptr(), T * C, C, n, T, C, self.proj_w, d, xp.data_ptr(), d, bias=self.proj_b,
                     out_batch_rows=xp.shape[1], out_row_offset=self.pos_pad)
        bn = self.pos_bn
        ops.gemm_raw(xp.data_ptr(), xp.shape[1] * d, d, n, T, self.pos_k * bn, self.pos_w, d, ws["tmp"].data_ptr(), d,
                     bias=self.pos_b, add_src_ptr=ws["x0"].data_ptr(), add_batch_rows=T, out_batch_rows=T,
                     flags=ops.GEMM_GELU | ops.GEMM_OUT_F32, conv_taps=self.pos_k, conv_stride=1,
                     a_rows_per_batch=xp.shape[1], a_col_per_ntile=bn, a_cols=d, force_bn=bn,
                     work=2.0 * n * T * d * self.pos_k * self.pos_cg)  # algorithmic: each output sees its group only
        x, xh, tmp = ws["x"], ws["x_f16"], ws["tmp"]
        self._tr("conv6", ws["feat"].view(n, T, C))
        self._tr("proj", ws["x0"].view(n, T, d))
        self._tr("posconv_sum", tmp.view(n, T, d))
        ops.layernorm(tmp, self.ln0_g, self.ln0_b, out_f16=xh, out_f32=x)
        self._tr("ln0", x.view(n, T, d))
        rel_bias = pos_bias = None
        if self.wavlm:
            if T <= self.REL_BIAS_MAX_T:
                rel_bias = self._rel_bias_for(T)
            else:
                pos_bias = self._pos_bias_for(T)
        for li, L in enumerate(self.layers):
            ops.linear(xh, L["wqkv"], L["bqkv"], out=ws["qkv"])
            if self.wavlm:
                ops.wavlm_gate(x, T, self.n_heads, L["gate_w"], L["gate_b"], L["gate_c"], ws["gate"])
                ops.attention(ws["qkv"], n, T, self.n_heads, gate=ws["gate"], pos_bias=pos_bias, rel_bias=rel_bias,
                              out=ws["att"])
            else:
                ops.attention(ws["qkv"], n, T, self.n_heads, out=ws["att"])
            ops.linear(ws["att"], L["wo"], L["bo"], add_src=x, out=tmp)
            ops.layernorm(tmp, L["ln1_g"], L["ln1_b"], out_f16=xh, out_f32=x)
            ops.linear(xh, L["w1"], L["b1"], gelu=True, out=ws["h1"])
            ops.linear(ws["h1"], L["w2"], L["b2"], add_src=x, out=tmp)
            ops.layernorm(tmp, L["ln2_g"], L["ln2_b"], out_f16=xh, out_f32=x)
            self._tr(f"layer{li}", x.view(n, T, d))
        keep = T if n_keep is None else min(n_keep, T)
        ops.heads(x.view(n, T, d), self.head_w, self.head_b, logits, frame_offset, step_frames, keep, frame_offsets=frame_offsets)

    def forward_waveforms(self, wav: torch.Tensor) -> torch.Tensor:
        """Drop-in ``model.forward``: (B, n_samples) fp32 -> (B, T, 1, C) fp32 logits."""
        wav = wav.to(self.device, torch.float32).contiguous()
        B, L = wav.shape
        T = conv_lengths(L)[-1]
        logits = torch.empty((B * T, len(self.labels)), dtype=torch.float32, device=self.device)
        self.forward_pcm(wav.reshape(-1), 0, B, L, L, logits, 0, T, T)
        return logits.view(B, T, 1, len(self.labels))
