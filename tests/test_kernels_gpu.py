"""Per-kernel parity on the GPU: every C-ABI entry point against the oracle / a torch fp32 restatement
of the same op on the same seeded inputs.  Integer / boolean outputs are bit-exact; floating-point
tolerances are stated per test."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import segma_oracle as O
from segma_b200 import ops, synth

pytestmark = pytest.mark.gpu


LABELS = synth.DEFAULT_LABELS


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g) * scale


def _close(got, ref, rtol, atol, what=""):
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs()
    bound = atol + rtol * ref.abs()
    bad = err > bound
    assert not bad.any(), f"{what}: {int(bad.sum())} / {bad.numel()} out of tolerance, max err {err.max():.4g} (ref max {ref.abs().max():.4g})"


# ---- GEMM -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,K,N", [(128, 64, 128), (300, 768, 768), (1500, 768, 2304), (257, 3072, 768), (199, 256, 1024),
                                   (1000, 128, 384), (130, 512, 192)])
def test_gemm_plain(cuda, M, K, N):
    a = _rand((M, K), 1).to(cuda, torch.float16)
    w = _rand((N, K), 2, K**-0.5).to(cuda, torch.float16)
    bias = _rand((N,), 3).to(cuda)
    out = ops.linear(a, w, bias)
    ref = a.float() @ w.float().T + bias
    # fp16 operands and output: 2^-11 relative rounding
    _close(out, ref, 2e-3, 2e-3, f"gemm {M}x{K}x{N}")


def test_gemm_epilogues(cuda):
    M, K, N = 700, 256, 512
    a = _rand((M, K), 4).to(cuda, torch.float16)
    w = _rand((N, K), 5, K**-0.5).to(cuda, torch.float16)
    bias = _rand((N,), 6).to(cuda)
    ref = a.float() @ w.float().T + bias
    out = ops.linear(a, w, bias, gelu=True)
    _close(out, F.gelu(ref), 2e-3, 2e-3, "gelu epilogue")
    out32 = ops.linear(a, w, bias, out_f32=True)
    _close(out32, ref, 1e-4, 1e-4, "fp32 out")
    gelu32 = ops.linear(a, w, bias, gelu=True, out_f32=True)
    _close(gelu32, F.gelu(ref), 1e-4, 1e-4, "gelu fp32 (A&S erf vs libm erf)")
    # residual update in place
    x = _rand((M, N), 7).to(cuda)
    want = x + ref
    ops.linear(a, w, bias, add_src=x, out=x)
    _close(x, want, 1e-4, 1e-4, "residual in place")
    # one table shared by every batch (position embedding), added after GELU; rows blocked 7 x 100
    pos = _rand((100, N), 8).to(cuda)
    out = torch.empty((M, N), device=cuda)
    flags = ops.GEMM_GELU | ops.GEMM_OUT_F32
    ops.gemm_raw(a.data_ptr(), 100 * K, K, 7, 100, K, w, N, out.data_ptr(), N, bias=bias, add_src_ptr=pos.data_ptr(),
                 add_batch_rows=0, flags=flags)
    idx = torch.arange(M, device=cuda) % 100
    _close(out, F.gelu(ref) + pos[idx], 1e-4, 1e-4, "gelu + shared table")
    # first 60 rows of every 100-row block only (last-layer row pruning), residual in place
    x2 = _rand((M, N), 9).to(cuda)
    want2 = x2.clone()
    sel = (torch.arange(M, device=cuda) % 100) < 60
    want2[sel] = (x2 + ref)[sel]
    ops.linear_rows(a, 7, 100, 60, w, bias, x2, add_src=x2)
    _close(x2, want2, 1e-4, 1e-4, "row-pruned residual")
    # no bias
    out = ops.linear(a, w, None, out_f32=True)
    _close(out, a.float() @ w.float().T, 1e-4, 1e-4, "no bias")


@pytest.mark.parametrize("C,N,taps,stride,T_in,pad", [(80, 128, 3, 1, 3000, 1), (128, 128, 3, 2, 3000, 1), (768, 768, 3, 2, 3000, 1),
                                                      (64, 64, 3, 2, 12799, 0), (64, 64, 2, 2, 399, 0)])
def test_gemm_conv(cuda, C, N, taps, stride, T_in, pad):
    B = 3
    x = _rand((B, C, T_in), 9)
    wt = _rand((N, C, taps), 10, (C * taps) ** -0.5)
    bias = _rand((N,), 11)
    xb, wb = x.to(torch.float16).float(), wt.to(torch.float16).float()
    ref = F.gelu(F.conv1d(xb, wb, bias, stride=stride, padding=pad)).transpose(1, 2)  # (B, T_out, N)
    T_out = ref.shape[1]
    rows_in = T_in + 2 * pad
    rows_in += (-rows_in) % stride
    x_tm = torch.zeros((B, rows_in, C), dtype=torch.float16)
    x_tm[:, pad:pad + T_in] = x.transpose(1, 2).to(torch.float16)
    w_tm = wt.permute(0, 2, 1).reshape(N, taps * C).contiguous().to(torch.float16)
    out = ops.conv1d_tm(x_tm.to(cuda), w_tm.to(cuda), bias.to(cuda), taps, stride, T_out, gelu=True, out_f32=True)
    _close(out, ref, 2e-3, 2e-3, f"conv C={C} N={N} k={taps} s={stride}")


# ---- layernorm ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("d", [128, 384, 768])
def test_layernorm(cuda, d):
    rows, period, keep = 1000, 250, 199
    x = _rand((rows, d), 12, 2.0).to(cuda) + 0.5
    g = (1 + 0.1 * _rand((d,), 13)).to(cuda)
    b = (0.1 * _rand((d,), 14)).to(cuda)
    ref = F.layer_norm(x, (d,), g, b, 1e-5)
    ob = torch.empty((rows, d), dtype=torch.float16, device=cuda)
    of = torch.empty((rows, d), dtype=torch.float32, device=cuda)
    mix = torch.zeros((rows // period, keep, d), device=cuda)
    ops.layernorm(x, g, b, out_f16=ob, out_f32=of, mix=mix, period=period, n_keep=keep, w_in=0.25, w_out=0.0, mix_init=True)
    ops.layernorm(x, g, b, mix=mix, period=period, n_keep=keep, w_in=0.0, w_out=0.5)
    _close(of, ref, 1e-5, 1e-5, "layernorm fp32")
    _close(ob, ref, 1e-3, 1e-3, "layernorm fp16")
    want = 0.25 * x.view(-1, period, d)[:, :keep] + 0.5 * ref.view(-1, period, d)[:, :keep]
    _close(mix, want, 1e-5, 1e-5, "layer mix")
    # only the kept rows
    ok = torch.full((rows, d), 7.0, dtype=torch.float16, device=cuda)
    ops.layernorm(x, g, b, out_f16=ok, period=period, n_keep=keep, only_kept=True)
    okv = ok.view(-1, period, d)
    _close(okv[:, :keep], ref.view(-1, period, d)[:, :keep], 1e-3, 1e-3, "only_kept rows")
    assert (okv[:, keep:] == 7.0).all()


def test_cast_f16(cuda):
    x = _rand((77, 256), 15).to(cuda)
    dst = torch.empty((77, 256), dtype=torch.float16, device=cuda)
    ops.cast_f16(x, dst)
    assert torch.equal(dst, x.to(torch.float16))


def test_linear_fp16_output_saturates(cuda):
    """Results outside the fp16 range are stored as +-65504, not infinity (a deliberate, documented divergence from a
    plain fp16 cast: the fp32 reference holds a finite value there too)."""
    a = torch.full((128, 64), 100.0, dtype=torch.float16, device=cuda)
    w = torch.full((32, 64), 100.0, dtype=torch.float16, device=cuda)
    w[16:] = -100.0
    out = ops.linear(a, w)
    assert torch.isfinite(out).all()
    assert (out[:, :16] == 65504.0).all() and (out[:, 16:] == -65504.0).all()


# ---- attention -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("T,H,B,nq", [(199, 2, 3, 199), (1500, 2, 2, 1500), (1500, 12, 1, 199), (64, 1, 1, 64), (65, 3, 2, 65),
                                      (1, 1, 1, 1), (31, 2, 1, 31), (32, 1, 2, 32), (33, 1, 1, 20), (129, 2, 1, 129)])
def test_attention(cuda, T, H, B, nq):
    d = H * 64
    qkv = _rand((B * T, 3 * d), 16).to(torch.float16)
    qkv[:, :d] *= 0.35
    out = ops.attention(qkv.to(cuda), B, T, H, n_query=nq)
    q, k, v = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = (torch.softmax(q @ k.transpose(-1, -2), -1) @ v).permute(0, 2, 1, 3).reshape(B, T, d)
    got = out.view(B, T, d)[:, :nq]
    _close(got, ref[:, :nq], 3e-3, 3e-3, f"attention T={T}")


@pytest.mark.parametrize("T", [199, 1500])
def test_attention_reference_max_rescale(cuda, T):
    """Scores that keep growing along the key axis (every later chunk beats the running reference max by far more than
    the 2^10 laziness bound, up to values whose exponentials overflow fp16) exercise the in-place rescale of O and l."""
    H, B = 2, 1
    d = H * 64
    g = torch.Generator().manual_seed(31)
    qkv = torch.randn((B * T, 3 * d), generator=g)
    ramp = torch.linspace(-1.0, 1.0, T)[:, None]
    qkv[:, :d] = 0.9 + 0.05 * qkv[:, :d]                       # q: almost constant direction
    qkv[:, d:2 * d] = ramp * 1.2 + 0.05 * qkv[:, d:2 * d]     # k: grows with the key index -> scores span about +-70
    qkv = qkv.to(torch.float16)
    out = ops.attention(qkv.to(cuda), B, T, H)
    q, k, v = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    s_ = q @ k.transpose(-1, -2)
    assert s_.max() - s_.min() > 100
    ref = (torch.softmax(s_, -1) @ v).permute(0, 2, 1, 3).reshape(B * T, d)
    _close(out, ref, 3e-3, 3e-3, f"attention with growing scores T={T}")


@pytest.mark.parametrize("T,padded", [(199, False), (199, True), (300, True), (64, True)])
def test_attention_wavlm_bias(cuda, T, padded):
    """padded=True: bias rows at a 16-byte pitch, what the one (tcgen05) kernel reads with 128-bit loads;
    False: an unpadded table with 796-byte rows is rejected -- there is no second backend to fall to."""
    H, B = 2, 2
    d = H * 64
    qkv = _rand((B * T, 3 * d), 17).to(torch.float16)
    qkv[:, :d] *= 0.35
    gate = (1.0 + 0.3 * _rand((B, H, T), 18)).contiguous()
    pos = _rand((H, T, T), 19).contiguous()
    pos_dev = pos.to(cuda)
    if padded:
        ld = (T + 3) // 4 * 4
        buf = torch.zeros((H, T, ld), device=cuda)
        buf[:, :, :T] = pos_dev
        pos_dev = buf[:, :, :T]
    else:
        with pytest.raises(ops.SegmaNativeError, match="16-byte aligned"):
            ops.attention(qkv.to(cuda), B, T, H, gate=gate.to(cuda), pos_bias=pos_dev)
        return
    out = ops.attention(qkv.to(cuda), B, T, H, gate=gate.to(cuda), pos_bias=pos_dev)
    q, k, v = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) + gate[..., None] * pos[None]
    ref = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B * T, d)
    _close(out, ref, 3e-3, 3e-3, "attention + gated bias")


@pytest.mark.parametrize("T", [199, 64, 300, 33])
def test_attention_wavlm_toeplitz_bias(cuda, T):
    """segma_attention_rel against the explicit (H, T, T) table built from the same Toeplitz vector."""
    H, B = 3, 2
    d = H * 64
    qkv = _rand((B * T, 3 * d), 27).to(torch.float16)
    qkv[:, :d] *= 0.35
    gate = (1.0 + 0.3 * _rand((B, H, T), 28)).contiguous()
    rel = _rand((H, 2 * T - 1), 29).contiguous()
    idx = torch.arange(T)[None, :] - torch.arange(T)[:, None] + T - 1
    pos = rel[:, idx]  # (H, T, T)
    out = ops.attention(qkv.to(cuda), B, T, H, gate=gate.to(cuda), rel_bias=rel.to(cuda))
    q, k, v = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) + gate[..., None] * pos[None]
    ref = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B * T, d)
    _close(out, ref, 3e-3, 3e-3, "attention + Toeplitz gated bias")


# ---- LSTM + heads -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("H,dirs,S", [(128, 2, 9), (64, 1, 9), (256, 2, 5), (128, 2, 128), (64, 2, 128)])
def test_lstm_layer(cuda, H, dirs, S):
    """H = 64 / 128: ``lstm_layer_smem_kernel`` (weights on the SM); H = 256: ``lstm_layer_kernel`` (weights through
    L2).  S = 128 is the reference's default batch: the recurrence is as long as it gets (inference.py:486-491)."""
    N, D = 21, 96
    sd = {}
    synth._lstm(sd, "lstm_shared.", D, synth.LSTMDims(H, 1, dirs == 2), 3)
    x = _rand((S, N, D), 20)
    ref = O.lstm_seq_first(sd, x)
    pres, whh = [], []
    for suffix in ("", "_reverse")[:dirs]:
        w_ih, w_hh = sd[f"lstm_shared.weight_ih_l0{suffix}"], sd[f"lstm_shared.weight_hh_l0{suffix}"]
        b = sd[f"lstm_shared.bias_ih_l0{suffix}"] + sd[f"lstm_shared.bias_hh_l0{suffix}"]
        pres.append(x @ w_ih.T + b)
        whh.append(w_hh.T.contiguous())
    pre = torch.cat(pres, dim=-1).contiguous().to(cuda)
    out = ops.lstm_layer(pre, torch.stack(whh).contiguous().to(cuda), H)
    # W_hh keeps fp32 precision on the SM (hi + lo halves for H = 128); state and gates are fp32
    _close(out, ref, 2e-5, 2e-5, "lstm layer")


def test_lstm_stack_split_precision_projection(cuda):
    """``LstmHeads`` (2 bidirectional layers + heads) over a 128-step recurrence on inputs that resemble each other
    from step to step, the regime of a real file (a large component common to all windows): rounding W_ih or the
    input to fp16 would be a systematic perturbation that adds up along the recurrence; the split-precision
    projection keeps the whole tail at fp32 level."""
    from segma_b200.engine import LstmHeads

    S, N, D = 128, 37, 256
    sd = {}
    synth._lstm(sd, "lstm_shared.", D, synth.LSTMDims(128, 2, True), 11)
    synth._heads(sd, LABELS, 256, 11)
    common = _rand((1, N, D), 30) * 1.7
    x = (common + 0.3 * _rand((S, N, D), 31)).contiguous()
    ref = O._heads(sd, O.lstm_seq_first(sd, x), LABELS).reshape(S * N, 4)
    tail = LstmHeads(sd, LABELS, cuda)
    logits = torch.empty((S * N, 4), device=cuda)
    tail.run(x.reshape(S * N, D).to(cuda), S, N, logits, 0, N, N)
    err = (logits.cpu() - ref).abs().max().item()
    print(f"LSTM tail over 128 steps: max|err| {err:.3g}, logit std {ref.std().item():.3g}")
    assert err <= 1e-4 * max(1.0, ref.std().item())


def test_cast_f16_split(cuda):
    x = (_rand((37, 64), 40) * torch.logspace(-4, 2, 64)).contiguous()
    dst = torch.empty((37, 192), dtype=torch.float16, device=cuda)
    ops.cast_f16_split(x.to(cuda), dst)
    d = dst.cpu()
    hi = x.half()
    assert torch.equal(d[:, :64], hi) and torch.equal(d[:, 128:], hi)
    assert torch.equal(d[:, 64:128], (x - hi.float()).half())
    w = ops.split_weight(x)
    assert torch.equal(w[:, :64], hi) and torch.equal(w[:, 64:128], hi) and torch.equal(w[:, 128:], (x - hi.float()).half())


def test_heads(cuda):
    S, N, Fd, C, keep = 5, 30, 256, 4, 19
    feat = _rand((S, N, Fd), 21).to(cuda)
    w = _rand((C, Fd), 22, 0.1).to(cuda)
    b = _rand((C,), 23).to(cuda)
    logits = torch.full((7 + S * keep, C), float("nan"), device=cuda)
    ops.heads(feat, w, b, logits, 7, keep, keep)
    ref = (feat[:, :keep] @ w.T + b).reshape(-1, C)
    _close(logits[7:], ref, 1e-5, 1e-5, "heads")
    assert torch.isnan(logits[:7]).all()


# ---- log-mel -----------------------------------------------------------------------------------------------
def _logmel_close(got, ref, what):
    # |a-b| <= 1e-4 * max(1, |b|)
    err = (got.cpu() - ref).abs()
    bound = 1e-4 * torch.clamp(ref.abs(), min=1.0)
    assert (err <= bound).all(), f"{what}: max err {err.max():.3g}, {(err > bound).sum()} elements out"


def test_mel_filters_match_oracle(cuda):
    assert np.array_equal(ops.mel_filters(), O.whisper_mel_filters().astype(np.float32))


def test_logmel_windows(cuda):
    n = 63680 * 3 + 64000
    pcm = torch.from_numpy(synth.synth_audio(n, 5))
    f32, tm = ops.logmel(pcm.to(cuda), 4, 64000, 63680, out_f32=True, out_tm=True)
    for i in range(4):
        ref = O.whisper_logmel(pcm[i * 63680: i * 63680 + 64000])
        _logmel_close(f32[i], ref, f"window {i}")
        assert torch.equal(tm[i, 1:3001].float().cpu(), f32[i].T.to(torch.float16).float().cpu())
        assert (tm[i, 0] == 0).all() and (tm[i, 3001] == 0).all()


@pytest.mark.parametrize("L", [400, 1000, 33280, 63999])
def test_logmel_tail_and_silence(cuda, L):
    pcm = torch.from_numpy(synth.synth_audio(L, 6))
    f32, _ = ops.logmel(pcm.to(cuda), 1, L, 63680)
    _logmel_close(f32[0], O.whisper_logmel(pcm), f"tail L={L}")
    z = torch.zeros(64000)
    f32, _ = ops.logmel(z.to(cuda), 1, 64000, 63680)
    _logmel_close(f32[0], O.whisper_logmel(z), "all-zero window")


@pytest.mark.parametrize("step,win", [(1601, 64000), (4000, 32000), (63680, 64000)])
def test_logmel_many_windows_persistent_clusters(cuda, step, win):
    """More windows than resident clusters (every cluster walks several windows and finishes window i behind the
    transform of window i + 1), an odd window step (unaligned windows take the scalar staging path), and an audio
    buffer that ends inside the batch: partial windows, then windows without any sample."""
    n_w = 330 if step < 60000 else 40
    n = step * (n_w - 6) + win // 3
    pcm = torch.from_numpy(synth.synth_audio(n, 11))
    dev = pcm.to(cuda)
    f32, tm = ops.logmel(dev, n_w, win, step, out_f32=True, out_tm=True)
    picks = sorted({0, 1, 2, 141, 142, 143, 200, n_w - 8, n_w - 7, n_w - 6, n_w - 5, n_w - 1} & set(range(n_w)))
    for i in picks:
        chunk = pcm[i * step: i * step + win]
        if chunk.numel() >= 400:  # the oracle (like the reference) needs a full STFT frame
            _logmel_close(f32[i], O.whisper_logmel(chunk), f"window {i} of {n_w}")
        if chunk.numel() == 0:  # past the end of the audio: log10(1e-10) clamped and scaled
            assert (f32[i] == -1.5).all()
            continue
        alone, _ = ops.logmel(dev[i * step:], 1, win, step)
        assert torch.equal(alone[0], f32[i]), f"window {i}: batch position changes the result"
    assert torch.equal(tm[:, 1:3001].float(), f32.transpose(1, 2).to(torch.float16).float())
    assert (tm[:, 0] == 0).all() and (tm[:, 3001] == 0).all()
    only_f32, _ = ops.logmel(dev, n_w, win, step)  # the scratch rows are dropped from L2 on this path
    assert torch.equal(only_f32, f32)


def test_logmel_custom_filterbank(cuda):
    """`segma_logmel_set_filters`: a bank with filters wider than the 15 taps the kernel keeps in shared memory (those
    read their taps from global memory), an empty filter and single-bin filters; then back to the built-in bank."""
    default = ops.mel_filters()
    rng = np.random.default_rng(3)
    bank = np.zeros((201, 80), dtype=np.float32)
    for m in range(80):
        width = [1, 2, 7, 15, 16, 23, 32][m % 7]
        lo = int(rng.integers(0, 201 - width + 1))
        bank[lo:lo + width, m] = rng.uniform(0.01, 0.05, size=width).astype(np.float32)
    bank[:, 5] = 0.0  # a filter without any bin: log10 of the 1e-10 floor
    pcm = torch.from_numpy(synth.synth_audio(63680 + 64000, 9))
    try:
        ops.set_mel_filters(bank)
        assert np.array_equal(ops.mel_filters(), bank)
        f32, _ = ops.logmel(pcm.to(cuda), 2, 64000, 63680)
        for i in range(2):
            _logmel_close(f32[i], O.whisper_logmel(pcm[i * 63680: i * 63680 + 64000], bank), f"custom bank, window {i}")
        too_wide = bank.copy()
        too_wide[0:40, 7] = 0.01
        with pytest.raises(Exception):
            ops.set_mel_filters(too_wide)
    finally:
        ops.set_mel_filters(default)
    f32, _ = ops.logmel(pcm.to(cuda), 1, 64000, 63680)
    _logmel_close(f32[0], O.whisper_logmel(pcm[:64000]), "built-in bank restored")


# ---- stitch + decode ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("F_,sf,nw,tail", [(199, 199, 5, 103), (199, 100, 7, 0), (99, 10, 12, 37), (399, 40, 3, 399)])
def test_stitch(cuda, F_, sf, nw, tail):
    C = 4
    wins = [_rand((F_, C), 30 + i) for i in range(nw)]
    offs = [i * sf for i in range(nw)]
    if tail:
        wins.append(_rand((tail, C), 99))
        offs.append(nw * sf)
    n_frames = max(o + w.shape[0] for o, w in zip(offs, wins))
    ref = O.stitch_mean(wins, offs, n_frames)
    got = ops.stitch(torch.cat(wins).to(cuda), nw, F_, sf, tail, n_frames)
    if sf == F_:
        assert torch.equal(got.cpu(), ref)  # concatenation is exact
    else:
        _close(got, ref, 1e-6, 1e-6, "stitch")


@pytest.mark.parametrize("n,C,p", [(1, 4, 1.0), (5, 4, 0.5), (1024, 4, 0.5), (1025, 4, 0.3), (100_000, 4, 0.5), (4097, 3, 0.9),
                                   (3000, 7, 0.1), (2048, 4, 0.0), (2048, 4, 1.0), (5000, 12, 0.4), (1025, 9, 0.6)])
def test_decode_intervals_bit_exact(cuda, n, C, p):
    """C <= 8 runs the bit-plane kernels (``decode_plane_kernel``), more labels the word-per-frame kernels
    (``decode_count_kernel`` / ``decode_write_kernel``)."""
    rng = np.random.default_rng(n + C)
    # blocky logits so that runs of every length occur
    base = rng.standard_normal((n // 7 + 1, C)).repeat(7, axis=0)[:n] + 0.3 * rng.standard_normal((n, C))
    logits = torch.from_numpy((base + (2 * p - 1) * 3).astype(np.float32))
    thr = [0.5, 0.3, 0.7, 0.5, 0.45, 0.55, 0.6, 0.5, 0.35, 0.65, 0.5, 0.4][:C]
    mask_ref = O.apply_thresholds(logits, thr)
    mask = ops.threshold_mask(logits.to(cuda), thr)
    assert torch.equal(mask.cpu(), mask_ref)
    table = ops.decode_intervals(logits.to(cuda), thr).cpu().numpy()
    ref = O.interval_table(mask_ref.numpy(), C)
    assert table.shape[0] == ref.shape[0]
    assert np.array_equal(table[:, 1:], ref)
    assert (table[:, 0] == 0).all()


def test_decode_multi_file_and_overflow(cuda):
    rng = np.random.default_rng(1)
    lens = [1500, 0, 1, 1024, 5000]
    offs = np.concatenate([[0], np.cumsum(lens)])
    logits = torch.from_numpy(rng.standard_normal((int(offs[-1]), 4)).astype(np.float32))
    thr = [0.5] * 4
    table = ops.decode_intervals(logits.to(cuda), thr, file_offsets=offs, capacity=8).cpu().numpy()
    rows = []
    for f, (a, b) in enumerate(zip(offs[:-1], offs[1:])):
        t = O.interval_table(O.apply_thresholds(logits[a:b], thr).numpy(), 4)
        rows.append(np.concatenate([np.full((t.shape[0], 1), f), t], axis=1))
    assert np.array_equal(table, np.concatenate(rows))


@pytest.mark.parametrize("C", [1, 4, 5, 8])
def test_decode_multi_file_logit_mode_runs_across_blocks(cuda, C):
    """Logit-domain cuts (the product's mode) on several files of awkward lengths, with runs long enough to span the
    1024-frame blocks and the 32-frame words of the bit-plane kernels."""
    rng = np.random.default_rng(40 + C)
    lens = [1023, 1024, 1025, 0, 31, 32, 33, 7000, 2048, 1]
    offs = np.concatenate([[0], np.cumsum(lens)])
    n = int(offs[-1])
    base = rng.standard_normal((n // 211 + 1, C)).repeat(211, axis=0)[:n] + 0.2 * rng.standard_normal((n, C))
    logits = torch.from_numpy(base.astype(np.float32))
    cuts = [0.1 * c - 0.2 for c in range(C)]
    table = ops.decode_intervals(logits.to(cuda), cuts, file_offsets=offs, mode=ops.DECODE_LOGIT).cpu().numpy()
    rows = []
    for f, (a, b) in enumerate(zip(offs[:-1], offs[1:])):
        mask = (logits[a:b] > torch.tensor(cuts)).numpy()
        t = O.interval_table(mask, C)
        rows.append(np.concatenate([np.full((t.shape[0], 1), f), t], axis=1))
    assert np.array_equal(table, np.concatenate(rows))


def test_decode_logit_cut_mode(cuda):
    from segma_b200.thresholds import logit_cut

    thr = [0.5, 0.3, 0.7, 0.9]
    cuts = [logit_cut(t) for t in thr]
    g = torch.Generator().manual_seed(0)
    logits = torch.randn((50_000, 4), generator=g) * 2
    # plant values straddling the cut by one ulp
    for c, cut in enumerate(cuts):
        v = torch.tensor(cut)
        logits[c * 10 + 0, c] = v
        logits[c * 10 + 1, c] = torch.nextafter(v, torch.tensor(float("inf")))
        logits[c * 10 + 2, c] = torch.nextafter(v, torch.tensor(float("-inf")))
    ref = O.apply_thresholds(logits, thr)
    got = ops.threshold_mask(logits.to(cuda), cuts, mode=ops.DECODE_LOGIT)
    assert torch.equal(got.cpu(), ref)


# ---- wav2vec2 front end / WavLM gate / grouped positional conv -----------------------------------------------
@pytest.mark.parametrize("L,C,gain", [(64000, 128, 1.0), (33280, 512, 1.0), (400, 128, 1.0), (64000, 512, 1e-3), (16000, 128, 3e-5)])
def test_w2v2_layer0(cuda, L, C, gain):
    """``gain``: quiet recordings (down to -90 dB) must come out as well as loud ones -- GroupNorm rescales them to
    unit variance, so the statistics (fp64 second moments of the input) are what matters."""
    n_win, step = 3, 1000
    pcm = torch.from_numpy(synth.synth_audio(L + (n_win - 1) * step, 31)) * gain
    w = _rand((C, 1, 10), 32, 0.5)
    g, b = 1 + 0.1 * _rand((C,), 33), 0.1 * _rand((C,), 34)
    T0 = (L - 10) // 5 + 1
    rows = T0 + (T0 & 1)
    out = torch.full((n_win, rows, C), 7.0, dtype=torch.float16, device=cuda)
    ss = torch.empty((n_win, C, 2), device=cuda)
    ops.w2v2_layer0(pcm.to(cuda), n_win, L, step, w.reshape(C, 10).contiguous().to(cuda), g.to(cuda), b.to(cuda), ss, out)
    wins = torch.stack([pcm[i * step: i * step + L] for i in range(n_win)])
    y = F.conv1d(wins.unsqueeze(1), w, None, stride=5)
    ref = F.gelu(F.group_norm(y, C, g, b, 1e-5)).transpose(1, 2)
    _close(out[:, :T0], ref, 1e-3, 1e-3, "layer 0")  # fp16 output rounding (2^-11) dominates
    assert (out[:, T0:] == 0).all()


def test_w2v2_layer0_silence(cuda):
    """All-zero audio (the reference's io fixture): GroupNorm variance 0 -> output gelu(beta)."""
    C, L = 128, 64000
    w = _rand((C, 10), 35, 0.5).to(cuda)
    g, b = (1 + 0.1 * _rand((C,), 36)).to(cuda), (0.3 * _rand((C,), 37)).to(cuda)
    out = torch.empty((1, 12800, C), dtype=torch.float16, device=cuda)
    ops.w2v2_layer0(torch.zeros(L, device=cuda), 1, L, L, w, g, b, torch.empty((1, C, 2), device=cuda), out)
    _close(out[0, :12799], F.gelu(b).expand(12799, C), 1e-3, 1e-3, "silent layer 0")


def test_w2v2_layer0_keeps_nan(cuda):
    """A NaN sample poisons its window's GroupNorm statistics; the GELU must hand the NaN on (the reference does), not
    clamp it away, and the neighbouring window stays clean."""
    C, L = 128, 4000
    pcm = torch.from_numpy(synth.synth_audio(2 * L, 44))
    pcm[100] = float("nan")
    w = _rand((C, 10), 35, 0.5).to(cuda)
    g, b = (1 + 0.1 * _rand((C,), 36)).to(cuda), (0.3 * _rand((C,), 37)).to(cuda)
    T0 = (L - 10) // 5 + 1
    out = torch.zeros((2, T0 + (T0 & 1), C), dtype=torch.float16, device=cuda)
    ops.w2v2_layer0(pcm.to(cuda), 2, L, L, w, g, b, torch.empty((2, C, 2), device=cuda), out)
    assert torch.isnan(out[0, :T0]).all()
    assert torch.isfinite(out[1]).all()


@pytest.mark.parametrize("H,T", [(2, 199), (12, 199), (16, 50), (5, 33), (1, 7)])
def test_wavlm_gate(cuda, H, T):
    B = 2
    x = _rand((B * T, H * 64), 38).to(cuda)
    gw, gb, gc = _rand((8, 64), 39, 0.2).to(cuda), _rand((8,), 40, 0.2).to(cuda), (1 + 0.2 * _rand((H,), 41)).to(cuda)
    gate = torch.empty((B, H, T), device=cuda)
    ops.wavlm_gate(x, T, H, gw, gb, gc, gate)
    xh = x.view(B, T, H, 64).permute(0, 2, 1, 3)
    gab = torch.sigmoid((xh @ gw.T + gb).view(B, H, T, 2, 4).sum(-1))
    ref = gab[..., 0] * (gab[..., 1] * gc.view(1, H, 1) - 1.0) + 2.0
    _close(gate, ref, 1e-5, 1e-5, "wavlm gate")


@pytest.mark.parametrize("dims", [synth.W2V2_TEST, synth.HUBERT_BASE])
def test_grouped_positional_conv(cuda, dims):
    """The grouped pos-conv + GELU + residual through the engine's packed weights vs F.conv1d."""
    from segma_b200.engine_w2v2 import W2V2Engine

    sd = synth.hubert_hydra_state_dict(dims, seed=12)
    eng = W2V2Engine(sd, synth.DEFAULT_LABELS, device=cuda)
    n, T, d = 2, 199, dims.d_model
    x0 = _rand((n, T, d), 42).to(torch.float16)
    rows_p = (T + eng.pos_k + 7) // 8 * 8
    xp = torch.zeros((n, rows_p, d), dtype=torch.float16, device=cuda)
    xp[:, eng.pos_pad: eng.pos_pad + T] = x0.to(cuda)
    x0f = x0.float().to(cuda).reshape(n * T, d).contiguous()
    out = torch.empty((n * T, d), device=cuda)
    bn = eng.pos_bn
    ops.gemm_raw(xp.data_ptr(), rows_p * d, d, n, T, eng.pos_k * bn, eng.pos_w, d, out.data_ptr(), d, bias=eng.pos_b,
                 add_src_ptr=x0f.data_ptr(), add_batch_rows=T, out_batch_rows=T, flags=ops.GEMM_GELU | ops.GEMM_OUT_F32,
                 conv_taps=eng.pos_k, conv_stride=1, a_rows_per_batch=rows_p, a_col_per_ntile=bn, a_cols=d, force_bn=bn)
    w = O._pos_conv_weight(sd, "wav2vec2.encoder.transformer.").to(torch.float16).float()
    pc = F.conv1d(x0.float().transpose(1, 2), w, sd["wav2vec2.encoder.transformer.pos_conv_embed.conv.bias"],
                  padding=eng.pos_k // 2, groups=dims.pos_groups)[..., :-1]
    ref = x0.float() + F.gelu(pc.transpose(1, 2))
    _close(out.view(n, T, d), ref, 3e-3, 3e-3, "positional conv")


# ---- extensions: hysteresis, gap merging, minimum duration --------------------------------
@pytest.mark.parametrize("n,C", [(1, 4), (1023, 4), (1024, 4), (5000, 3), (40_000, 8)])
def test_hysteresis_decode(cuda, n, C):
    g = torch.Generator().manual_seed(n)
    logits = torch.cumsum(torch.randn((n, C), generator=g) * 0.4, dim=0)
    logits = logits - logits.mean(0, keepdim=True)
    lo, hi = [-0.3 + 0.05 * c for c in range(C)], [0.4 + 0.05 * c for c in range(C)]
    mask = O.hysteresis_mask(logits, lo, hi)
    want = O.interval_table(mask, C)
    got = ops.decode_intervals(logits.to(cuda), lo, mode=ops.DECODE_LOGIT, onset=hi).cpu().numpy()
    assert np.array_equal(got[:, 1:], want)
    # onset == offset is the plain threshold rule
    same = ops.decode_intervals(logits.to(cuda), lo, mode=ops.DECODE_LOGIT, onset=lo).cpu().numpy()
    plain = ops.decode_intervals(logits.to(cuda), lo, mode=ops.DECODE_LOGIT).cpu().numpy()
    assert np.array_equal(same, plain)


def test_hysteresis_resets_at_file_boundaries(cuda):
    logits = torch.full((3000, 2), 0.1)  # between the cuts: keeps whatever state it has
    logits[10] = 5.0                     # file 0 switches on at frame 10 and stays on
    offs = [0, 1500, 3000]               # file 1 never crosses the onset: stays off
    got = ops.decode_intervals(logits.to(cuda), [0.0, 0.0], file_offsets=offs, mode=ops.DECODE_LOGIT, onset=[1.0, 1.0])
    assert got.cpu().tolist() == [[0, 0, 3200, 480000], [0, 1, 3200, 480000]]


def test_postprocess_intervals(cuda):
    rng = np.random.default_rng(3)
    rows = []
    for f in range(3):
        for c in range(4):
            t = 0
            for _ in range(int(rng.integers(0, 700))):
                t += int(rng.integers(1, 30)) * 320
                e = t + int(rng.integers(1, 40)) * 320
                rows.append((f, c, t, e))
                t = e
    table = np.array(rows, dtype=np.int32)
    for gap, dur in [(0, 0), (320 * 5, 0), (0, 320 * 10), (320 * 8, 320 * 20)]:
        got = ops.postprocess_intervals(torch.from_numpy(table).to(cuda), gap, dur).cpu().numpy()
        assert np.array_equal(got, O.postprocess_table(table, gap, dur)), (gap, dur)


def test_intervals_struct_vs_reference_golden(cuda):
    """``segma_b200.structs.Intervals`` (merge on the device) against what the reference's own ``Intervals`` produced
    for every insertion sequence of /root/reference/tests/test_interval.py -- after each ``add`` -- and for shuffled
    decode tables with overlapping / bridging extras (tests/golden/postprocess.json, oracle/make_golden.py)."""
    import json
    from pathlib import Path

    from segma_b200.structs import Intervals

    cases = json.loads((Path(__file__).parent / "golden" / "postprocess.json").read_text())["cases"]
    assert len(cases) >= 30
    for case in cases:
        adds = [tuple(r) for r in case["adds"]]
        if "after_each_add" in case:
            iv = Intervals()
            assert iv.intervals == [] and len(iv) == 0
            for item, state in zip(adds, case["after_each_add"]):
                iv.add(item)
                assert iv.intervals == [tuple(r) for r in state]
        iv = Intervals()
        iv.extend(adds)
        assert iv.intervals == [tuple(r) for r in case["final"]]
        assert list(iv) == iv.intervals and len(iv) == len(case["final"])


def test_postprocess_overlapping_rows_across_chunks(cuda):
    """Nested and overlapping rows (running maximum of the ends) over more rows than one scan chunk of 1024."""
    rng = np.random.default_rng(11)
    rows = []
    for f in range(2):
        for c in range(3):
            starts = np.sort(rng.integers(0, 400_000, size=int(rng.integers(900, 2600))))
            for s in starts:
                rows.append((f, c, int(s), int(s) + int(rng.integers(0, 900))))
    table = np.array(rows, dtype=np.int32)
    for gap, dur in [(0, 0), (50, 0), (0, 400), (120, 1000)]:
        got = ops.postprocess_intervals(torch.from_numpy(table).to(cuda), gap, dur).cpu().numpy()
        want = O.postprocess_table(table, gap, dur)
        assert np.array_equal(got, want), (gap, dur)
        assert 0 < want.shape[0] < table.shape[0]


def test_decode_long_input_parallel_scan(cuda):
    """More than 32 768 (block, label) counters: the scan of pass 2 runs as three parallel kernels."""
    n, C = 9_000_000, 4
    g = torch.Generator().manual_seed(7)
    base = torch.randn((n // 500 + 1, C), generator=g).repeat_interleave(500, dim=0)[:n]
    logits = (base + 0.02 * torch.randn((n, C), generator=g)).contiguous()
    offs = [0, 3_000_001, 3_000_001, 7_654_321, n]
    cuts = [0.1, -0.2, 0.0, 0.3]
    table = ops.decode_intervals(logits.to(cuda), cuts, file_offsets=offs, mode=ops.DECODE_LOGIT).cpu().numpy()
    rows = []
    for f, (a, b) in enumerate(zip(offs[:-1], offs[1:])):
        mask = (logits[a:b] > torch.tensor(cuts)).numpy()
        t = O.interval_table(mask, C)
        rows.append(np.concatenate([np.full((t.shape[0], 1), f), t], axis=1))
    assert np.array_equal(table, np.concatenate(rows))
