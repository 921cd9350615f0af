"""Device-side execution of one forward call of the reference loop (a batch of windows of one file).

``WhisperEngine`` runs what ``SurgicalHydra.forward`` / ``HydraWhisper.forward`` compute
(/root/reference/src/segma/models/whisper/surgical_hydra.py:80-109, hydra.py:71-87) fused with the
windowing and the log-mel hook in front of them (inference.py:148-152, hydra.py:197-201): it takes raw
PCM resident on the device and writes frame logits straight onto the file timeline.  Every arithmetic
step is a libsegma_b200 kernel; torch only owns the buffers and the stream.

Weights are packed once from the reference ``state_dict``: fp16 GEMM operands
(q/k/v fused, the query scale 64**-0.5 folded into W_q/b_q -- exact, it is a power of two --, conv
kernels re-laid tap-major), fp32 biases / LayerNorm / LSTM recurrence / heads.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import ops

HEAD_DIM = 64
N_CTX = 1500
MEL_FRAMES = 3000


def _f16(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.float16).contiguous()


def _f32(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def resolve_device(device) -> torch.device:
    """``"cuda"`` / ``cuda:N`` -> a device with an explicit index (the current one when none is given)."""
    dev = torch.device(device)
    if dev.type == "cuda" and dev.index is None and torch.cuda.is_available():
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def move_to(obj, device):
    """Packed weights (tensors nested in lists / dicts / small holder objects) -> ``device``."""
    if isinstance(obj, torch.Tensor):
        return obj.to(device)
    if isinstance(obj, list):
        return [move_to(v, device) for v in obj]
    if isinstance(obj, tuple):
        return tuple(move_to(v, device) for v in obj)
    if isinstance(obj, dict):
        return {k: move_to(v, device) for k, v in obj.items()}
    if isinstance(obj, (_LstmLayer, LstmHeads)):
        for k, v in vars(obj).items():
            setattr(obj, k, move_to(v, device))
    return obj


@dataclass
class _LstmLayer:
    w_ih: torch.Tensor  # (n_dirs*4H, 3*in) fp16, split precision [W_hi | W_hi | W_lo]
    bias: torch.Tensor  # (n_dirs*4H) fp32 = b_ih + b_hh
    w_hh_t: torch.Tensor  # (n_dirs, H, 4H) fp32


class LstmHeads:
    """``lstm_shared`` (sequence axis = windows of the batch) followed by the per-label heads."""

    def __init__(self, sd: dict, labels, device, prefix: str = "lstm_shared."):
        self.layers: list[_LstmLayer] = []
        layer = 0
        while f"{prefix}weight_ih_l{layer}" in sd:
            sufs = [s for s in ("", "_reverse") if f"{prefix}weight_ih_l{layer}{s}" in sd]
            w_ih = torch.cat([sd[f"{prefix}weight_ih_l{layer}{s}"] for s in sufs], dim=0)
            bias = torch.cat([sd[f"{prefix}bias_ih_l{layer}{s}"] + sd[f"{prefix}bias_hh_l{layer}{s}"] for s in sufs])
            w_hh_t = torch.stack([sd[f"{prefix}weight_hh_l{layer}{s}"].T.contiguous() for s in sufs])
            # The recurrence runs over up to batch_size windows and the windows of a file resemble each other, so a
            # rounded W_ih / input is a *systematic* perturbation that adds up along it: the input projection runs at
            # split precision (hi + lo fp16 halves of both operands, one GEMM over K = 3 * in).
            self.layers.append(_LstmLayer(ops.split_weight(w_ih).to(device), _f32(bias, device), _f32(w_hh_t, device)))
            layer += 1
        self.hidden = self.layers[0].w_hh_t.shape[1] if self.layers else 0
        self.n_dirs = self.layers[0].w_hh_t.shape[0] if self.layers else 0
        self.head_w = _f32(torch.cat([sd[f"task_heads.linear_head_{lab}.weight"] for lab in labels], dim=0), device)
        self.head_b = _f32(torch.cat([sd[f"task_heads.linear_head_{lab}.bias"] for lab in labels], dim=0), device)
        self.n_labels = len(labels)

    def run(self, feat_f32: torch.Tensor, n_steps: int, n_rows: int, logits: torch.Tensor,
            frame_offset: int, step_frames: int, n_keep: int) -> None:
        """feat_f32 (n_steps*n_rows, F): LSTM input; each layer's input is split into fp16 hi / lo halves for the
        tensor-core projection."""
        cur = feat_f32
        for lay in self.layers:
            rows, feat = cur.shape
            a = torch.empty((rows, 3 * feat), dtype=torch.float16, device=cur.device)
            ops.cast_f16_split(cur, a)
            pre = ops.linear(a, lay.w_ih, lay.bias, out_f32=True)
            cur = ops.lstm_layer(pre.view(n_steps, n_rows, -1), lay.w_hh_t, self.hidden).view(rows, -1)
        ops.heads(cur.view(n_steps, n_rows, -1), self.head_w, self.head_b, logits, frame_offset, step_frames, n_keep)


class WhisperEngine:
    def __init__(self, sd: dict, labels, *, kind: str = "surgical_hydra", encoder_layers=None,
                 reduction: str = "weighted", n_keep: int = 199, device="cuda", prefix: str = "w_encoder."):
        self.device = resolve_device(device)
        with torch.cuda.device(self.device):
            ops.device_check()
        self.kind = kind
        self.labels = tuple(labels)
        self.n_keep = n_keep
        dev = self.device
        p = prefix
        d = sd[p + "conv1.weight"].shape[0]
        self.d = d
        self.n_heads = d // HEAD_DIM
        self.n_mels = sd[p + "conv1.weight"].shape[1]
        assert self.n_mels == 80 and d % 128 == 0, "Whisper encoder dims must be 80 mel bins and d_model % 128 == 0"
        # conv stem, tap-major K
        self.conv1_w = _f16(sd[p + "conv1.weight"].permute(0, 2, 1).reshape(d, -1), dev)
        self.conv1_b = _f32(sd[p + "conv1.bias"], dev)
        self.conv2_w = _f16(sd[p + "conv2.weight"].permute(0, 2, 1).reshape(d, -1), dev)
        self.conv2_b = _f32(sd[p + "conv2.bias"], dev)
        self.pos = _f32(sd[p + "embed_positions.weight"], dev)
        assert self.pos.shape == (N_CTX, d)
        self.layers = []
        i = 0
        scale = HEAD_DIM**-0.5
        while f"{p}layers.{i}.fc1.weight" in sd:
            lp = f"{p}layers.{i}."
            wq, bq = sd[lp + "self_attn.q_proj.weight"] * scale, sd[lp + "self_attn.q_proj.bias"] * scale
            wk = sd[lp + "self_attn.k_proj.weight"]
            wv, bv = sd[lp + "self_attn.v_proj.weight"], sd[lp + "self_attn.v_proj.bias"]
            self.layers.append(dict(
                ln1_g=_f32(sd[lp + "self_attn_layer_norm.weight"], dev), ln1_b=_f32(sd[lp + "self_attn_layer_norm.bias"], dev),
                wqkv=_f16(torch.cat([wq, wk, wv], dim=0), dev),
                bqkv=_f32(torch.cat([bq, torch.zeros_like(bq), bv]), dev),
                wo=_f16(sd[lp + "self_attn.out_proj.weight"], dev), bo=_f32(sd[lp + "self_attn.out_proj.bias"], dev),
                ln2_g=_f32(sd[lp + "final_layer_norm.weight"], dev), ln2_b=_f32(sd[lp + "final_layer_norm.bias"], dev),
                w1=_f16(sd[lp + "fc1.weight"], dev), b1=_f32(sd[lp + "fc1.bias"], dev),
                w2=_f16(sd[lp + "fc2.weight"], dev), b2=_f32(sd[lp + "fc2.bias"], dev),
            ))
            i += 1
        self.n_layers = i
        self.ffn = self.layers[0]["w1"].shape[0]
        self.lnf_g = _f32(sd[p + "layer_norm.weight"], dev)
        self.lnf_b = _f32(sd[p + "layer_norm.bias"], dev)
        # layer mix: weight of hidden_states[1:][i]; i = n_layers-1 is LN_f(h_N)
        self.mix_w = [0.0] * self.n_layers
        if kind == "surgical_hydra":
            use = list(range(self.n_layers)) if not encoder_layers else sorted(j - 1 for j in encoder_layers)
            w = sd["layer_weights"].detach().float().cpu()
            w = torch.softmax(w, dim=0) if reduction == "weighted" else w
            assert len(use) == w.numel(), "layer_weights does not match encoder_layers"
            for j, li in enumerate(use):
                self.mix_w[li] += float(w[j])
        elif kind == "hydra_whisper":
            self.mix_w[-1] = 1.0
        else:
            raise ValueError(f"unknown Whisper-family model kind '{kind}'")
        self.tail = LstmHeads(sd, labels, dev)
        self._slots: dict[int, tuple[int, dict]] = {}  # workspace slot -> (capacity in windows, buffers)
        self._ws: dict[str, torch.Tensor] = {}

    def to(self, device) -> "WhisperEngine":
        """Move the packed weights to another CUDA device; workspaces are re-created there on demand."""
        device = resolve_device(device)
        if device != self.device:
            for k, v in vars(self).items():
                if k not in ("_slots", "_ws", "device"):
                    setattr(self, k, move_to(v, device))
            self._slots, self._ws, self.device = {}, {}, device
            with torch.cuda.device(device):
                ops.device_check()
        return self

    # ---- workspace -------------------------------------------------------------------------------
    def _reserve(self, n: int, slot: int = 0) -> None:
        """Make ``self._ws`` the buffers of ``slot`` (one slot per concurrently running batch), grown to n windows."""
        cap, ws = self._slots.get(slot, (0, {}))
        if n <= cap:
            self._ws = ws
            return
        dev, d = self.device, self.d
        M = n * N_CTX
        ws = {}
        ws["mel_tm"] = torch.zeros((n, MEL_FRAMES + 2, 80), dtype=torch.float16, device=dev)
        ws["c1"] = torch.zeros((n, MEL_FRAMES + 2, d), dtype=torch.float16, device=dev)  # pad rows stay zero
        ws["x"] = torch.empty((M, d), dtype=torch.float32, device=dev)
        ws["xn"] = torch.empty((M, d), dtype=torch.float16, device=dev)
        ws["qkv"] = torch.empty((M, 3 * d), dtype=torch.float16, device=dev)
        ws["att"] = torch.empty((M, d), dtype=torch.float16, device=dev)
        ws["h1"] = torch.empty((M, self.ffn), dtype=torch.float16, device=dev)
        ws["mix"] = torch.empty((n, self.n_keep, d), dtype=torch.float32, device=dev)
        ws["mel_scratch"] = torch.empty(max(ops.logmel_scratch_bytes(n, 64_000 * 2), 1), dtype=torch.uint8, device=dev)
        self._slots[slot] = (n, ws)
        self._ws = ws

    # ---- stages ------------------------------------------------------------------------------------
    def encode_tm(self, mel_tm: torch.Tensor, n: int) -> None:
        """(n, 3002, 80) fp16 padded time-major log-mel -> ws['mix'] (n, n_keep, d) fp32."""
        ws, d, T = self._ws, self.d, N_CTX
        M = n * T
        x, xn, qkv, att, h1, mix = ws["x"][:M], ws["xn"][:M], ws["qkv"][:M], ws["att"][:M], ws["h1"][:M], ws["mix"][:n]
        ops.conv1d_tm(mel_tm[:n], self.conv1_w, self.conv1_b, 3, 1, MEL_FRAMES, gelu=True, out=ws["c1"][:n],
                      out_batch_rows=MEL_FRAMES + 2, out_row_offset=1)
        ops.conv1d_tm(ws["c1"][:n], self.conv2_w, self.conv2_b, 3, 2, T, gelu=True, add_src=self.pos, add_batch_rows=0,
                      out=x.view(n, T, d))
        mixed = False
        keep = self.n_keep
        for li, L in enumerate(self.layers):
            w_in = self.mix_w[li - 1] if li > 0 else 0.0
            if w_in != 0.0:
                ops.layernorm(x, L["ln1_g"], L["ln1_b"], out_f16=xn, mix=mix, period=T, n_keep=keep, w_in=w_in,
                              mix_init=not mixed)
                mixed = True
            else:
                ops.layernorm(x, L["ln1_g"], L["ln1_b"], out_f16=xn)
            ops.linear(xn, L["wqkv"], L["bqkv"], out=qkv)
            if li < self.n_layers - 1 or keep >= T:
                ops.attention(qkv, n, T, self.n_heads, out=att)
                ops.linear(att, L["wo"], L["bo"], add_src=x, out=x)
                ops.layernorm(x, L["ln2_g"], L["ln2_b"], out_f16=xn)
                ops.linear(xn, L["w1"], L["b1"], gelu=True, out=h1)
                ops.linear(h1, L["w2"], L["b2"], add_src=x, out=x)
            else:
                # last layer: nothing downstream reads positions >= n_keep, so only those query rows are
                # attended, projected and fed through the MLP (keys / values still span all T positions)
                ops.attention(qkv, n, T, self.n_heads, n_query=keep, out=att)
                ops.linear_rows(att, n, T, keep, L["wo"], L["bo"], x, add_src=x)
                ops.layernorm(x, L["ln2_g"], L["ln2_b"], out_f16=xn, period=T, n_keep=keep, only_kept=True)
                ops.linear_rows(xn, n, T, keep, L["w1"], L["b1"], h1, gelu=True)
                ops.linear_rows(h1, n, T, keep, L["w2"], L["b2"], x, add_src=x)
        ops.layernorm(x, self.lnf_g, self.lnf_b, mix=mix, period=T, n_keep=keep, w_in=0.0,
                      w_out=self.mix_w[-1], mix_init=not mixed, only_kept=True)

    def forward_pcm(self, pcm: torch.Tensor, start: int, n: int, win_len: int, step: int, logits: torch.Tensor,
                    frame_offset: int, step_frames: int, n_keep: int | None = None, slot: int = 0) -> None:
        """Windows ``pcm[start + i*step : +win_len]``, i < n, of a device-resident 1-D fp32 signal ->
        ``logits[(frame_offset + i*step_frames + r), :]`` for r < n_keep.  ``slot`` selects the workspace
        (batches running concurrently on different streams use different slots)."""
        with torch.cuda.device(self.device):
            self._reserve(n, slot)
            ws = self._ws
            need = ops.logmel_scratch_bytes(n, win_len)
            if ws["mel_scratch"].numel() < need:
                ws["mel_scratch"] = torch.empty(need, dtype=torch.uint8, device=self.device)
            view = pcm[start:]
            ops.logmel_into(view, n, win_len, step, ws["mel_tm"], ws["mel_scratch"])
            self.encode_tm(ws["mel_tm"], n)
            keep = self.n_keep if n_keep is None else n_keep
            self.tail.run(ws["mix"][:n].view(n * self.n_keep, self.d), n, self.n_keep, logits, frame_offset, step_frames,
                          keep)

    def forward_features(self, feats: torch.Tensor) -> torch.Tensor:
        """Drop-in ``model.forward``: (B, 80, 3000) fp32 log-mel -> (B, n_keep, 1, C) fp32 logits."""
        n = feats.shape[0]
        with torch.cuda.device(self.device):
            self._reserve(n)
            ws = self._ws
            tm = ws["mel_tm"][:n]
            tm[:, 1:-1] = feats.to(self.device).transpose(1, 2).to(torch.float16)  # layout change only
            self.encode_tm(ws["mel_tm"], n)
            logits = torch.empty((n * self.n_keep, len(self.labels)), dtype=torch.float32, device=self.device)
            self.tail.run(ws["mix"][:n].view(n * self.n_keep, self.d), n, self.n_keep, logits, 0, self.n_keep, self.n_keep)
        return logits.view(n, self.n_keep, 1, len(self.labels))
