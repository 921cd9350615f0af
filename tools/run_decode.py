"""Runs stitch + decode alone on a large batch (for ncu captures): python tools/run_decode.py [hours]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from segma_b200 import ops  # noqa: E402
from segma_b200.thresholds import logit_cut  # noqa: E402

hours = int(sys.argv[1]) if len(sys.argv) > 1 else 200
n = 180_000 * hours
offs = [i * 180_000 for i in range(hours + 1)]
cuts = [logit_cut(0.5)] * 4
big = (torch.randn((n // 50, 4), device="cuda").repeat_interleave(50, dim=0) + 0.05 * torch.randn((n, 4), device="cuda")).contiguous()
t = ops.decode_intervals(big, cuts, file_offsets=offs, mode=ops.DECODE_LOGIT)
n_iv = t.shape[0]
ops.stats.reset()
ops.stats.profile = True
for _ in range(5):
    ops.decode_intervals(big, cuts, file_offsets=offs, mode=ops.DECODE_LOGIT, capacity=n_iv)
torch.cuda.synchronize()
ms = min(e0.elapsed_time(e1) for nm, e0, e1, _ in ops.stats.events)
print(f"{hours} h, {n_iv} intervals: {ms:.3f} ms, {(16 * n + 16 * n_iv) / ms / 1e6:.0f} GB/s algorithmic")
