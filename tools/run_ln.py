"""Times the LayerNorm kernel alone: python tools/run_ln.py [rows]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from segma_b200 import ops  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 192000
d = 768
x = torch.randn((M, d), device="cuda")
g = torch.randn(d, device="cuda")
b = torch.randn(d, device="cuda")
o16 = torch.empty((M, d), dtype=torch.float16, device="cuda")
for _ in range(3):
    ops.layernorm(x, g, b, out_f16=o16)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.layernorm(x, g, b, out_f16=o16)
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 20 * 1e-3
print(f"{M} rows: {t * 1e6:.1f} us, {M * d * 6 / t / 1e9:.0f} GB/s")
