"""Runs the attention kernel alone (for ncu captures): python tools/run_attention.py [n_windows] [T]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from segma_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
H = 12
qkv = (torch.randn((n * T, 3 * H * 64), device="cuda") * 0.5).to(torch.float16)
out = torch.empty((n * T, H * 64), dtype=torch.float16, device="cuda")
for _ in range(3):
    ops.attention(qkv, n, T, H, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.attention(qkv, n, T, H, out=out)
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 10 * 1e-3
fl = 4.0 * T * T * 64 * H * n
print(f"{n} windows T={T}: burst {t * 1e3:.3f} ms, {fl / t / 1e12:.1f} TFLOP/s")
# steady state under the power cap: the clock governor needs a few hundred ms to settle
iters = max(10, int(0.8 / t))
for _ in range(iters):
    ops.attention(qkv, n, T, H, out=out)
torch.cuda.synchronize()
e0.record()
for _ in range(iters):
    ops.attention(qkv, n, T, H, out=out)
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / iters * 1e-3
print(f"{n} windows T={T}: sustained {t * 1e3:.3f} ms, {fl / t / 1e12:.1f} TFLOP/s ({iters} iterations)")
