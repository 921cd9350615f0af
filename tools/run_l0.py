"""Runs the wav2vec2 layer-0 front end (conv k10 s5 + GroupNorm + GELU) and the WavLM gate alone:
python tools/run_l0.py [n_windows]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from segma_b200 import ops, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
C, L, step = 512, 64000, 63680
T0 = (L - 10) // 5 + 1
pcm = torch.from_numpy(synth.synth_audio(step * (n - 1) + L, 0)).cuda()
g = torch.Generator(device="cuda").manual_seed(0)
w = torch.randn((C, 10), device="cuda", generator=g) * 0.3
gamma = 1.0 + 0.1 * torch.randn(C, device="cuda", generator=g)
beta = 0.1 * torch.randn(C, device="cuda", generator=g)
ss = torch.empty((n, C, 2), device="cuda")
out = torch.empty((n, T0 + 1, C), dtype=torch.float16, device="cuda")


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


t = timed(lambda: ops.w2v2_layer0(pcm, n, L, step, w, gamma, beta, ss, out))
print(f"layer0: {n} windows {t * 1e3:.3f} ms, {out.numel() * 2 / t / 1e9:.0f} GB/s written")
T, H = 199, 12
x = torch.randn((n * T, H * 64), device="cuda", generator=g)
gw = torch.randn((8, 64), device="cuda", generator=g) * 0.1
gb = torch.randn(8, device="cuda", generator=g)
gc = torch.randn(H, device="cuda", generator=g)
gate = torch.empty((n, H, T), device="cuda")
t = timed(lambda: ops.wavlm_gate(x, T, H, gw, gb, gc, gate))
print(f"gate: {n * T} rows {t * 1e6:.1f} us, {x.numel() * 4 / t / 1e9:.0f} GB/s read")
