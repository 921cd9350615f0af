"""Runs the log-mel front end alone (for ncu captures): python tools/run_logmel.py [n_windows]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from segma_b200 import ops, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
pcm = torch.from_numpy(synth.synth_audio(63680 * (n - 1) + 64000, 0)).cuda()
for _ in range(3):
    f32, _ = ops.logmel(pcm, n, 64000, 63680)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    f32, _ = ops.logmel(pcm, n, 64000, 63680)
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 10 * 1e-3
print(f"{n} windows: {t * 1e6 / n:.3f} us/window, {n * 1216000 / t / 1e9:.0f} GB/s algorithmic")
