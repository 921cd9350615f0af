"""Is the log-mel kernel's time a matter of the SM clock?  Times it cold, right after 3 s of tensor-core load (the
state bench.py's side measurements start from) and again after pauses, with the SM clock read through NVML."""
import sys
import time
from pathlib import Path

import pynvml
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from segma_b200 import ops, synth  # noqa: E402

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
n = 1024
pcm = torch.from_numpy(synth.synth_audio(63680 * (n - 1) + 64000, 0)).cuda()


def clock():
    return pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)


def run(tag, iters=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0 = clock()
    e0.record()
    for _ in range(iters):
        ops.logmel(pcm, n, 64000, 63680)
    e1.record()
    torch.cuda.synchronize()
    print(f"{tag}: {e0.elapsed_time(e1) / iters / n * 1e3:.3f} us/window, SM clock {c0} -> {clock()} MHz", flush=True)


for _ in range(3):
    ops.logmel(pcm, n, 64000, 63680)
torch.cuda.synchronize()
run("cold")
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
t0 = time.time()
while time.time() - t0 < 3.0:
    for _ in range(20):
        a @ a
    torch.cuda.synchronize()
run("right after 3 s of GEMMs")
run("again")
time.sleep(0.5)
run("after 0.5 s idle")
time.sleep(2.0)
run("after 2 s idle")
run("200 iterations", 200)
