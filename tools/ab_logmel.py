"""A/B of experimental builds of the log-mel kernel: python tools/ab_logmel.py lib1.so [lib2.so ...]
Each library is loaded through ctypes directly (segma_logmel only) and timed alternately on 1024 windows; the fp32
outputs of all libraries are compared with the first one's."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from segma_b200 import synth  # noqa: E402

n = 1024
pcm = torch.from_numpy(synth.synth_audio(63680 * (n - 1) + 64000, 0)).cuda()
libs = []
for path in sys.argv[1:]:
    lib = C.CDLL(path)
    lib.segma_logmel_scratch_bytes.restype = C.c_size_t
    lib.segma_logmel_scratch_bytes.argtypes = [C.c_int, C.c_int]
    lib.segma_logmel.restype = C.c_int
    lib.segma_logmel.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    libs.append((path, lib))
out = torch.empty((n, 80, 3000), device="cuda")
st = torch.cuda.current_stream().cuda_stream


def logmel_f64(x, mel):
    """Whisper log-mel of one window in float64 (numpy): the ground truth both the oracle's fp32 torch.stft and the
    kernel approximate."""
    import numpy as np
    x = np.concatenate([x.astype(np.float64), np.zeros(480000 - len(x))])
    xp = np.pad(x, 200, mode="reflect")
    hann = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(400) / 400)
    idx = np.arange(3000)[:, None] * 160 + np.arange(400)[None, :]
    spec = np.fft.rfft(xp[idx] * hann, axis=1)
    p = (spec.real ** 2 + spec.imag ** 2) @ mel.astype(np.float64)
    lg = np.log10(np.maximum(p, 1e-10))
    lg = np.maximum(lg, lg.max() - 8.0)
    return ((lg + 4.0) / 4.0).T


import numpy as np
from segma_b200 import ops  # noqa: E402
mel = ops.mel_filters()
truth_idx = [0, 1, 500, 1023]
truth = [logmel_f64(pcm[i * 63680: i * 63680 + 64000].cpu().numpy(), mel) for i in truth_idx]
ref = None
for rnd in range(3):
    for path, lib in libs:
        scratch = torch.empty(lib.segma_logmel_scratch_bytes(n, 64000), dtype=torch.uint8, device="cuda")
        call = lambda: lib.segma_logmel(pcm.data_ptr(), pcm.numel(), n, 64000, 63680, out.data_ptr(), None, scratch.data_ptr(), st)  # noqa: E731
        for _ in range(3):
            assert call() == 0
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(40):
            call()
        e1.record()
        torch.cuda.synchronize()
        if ref is None:
            ref = out.clone()
        worst = 0.0
        for i, t in zip(truth_idx, truth):
            got = out[i].double().cpu().numpy()
            worst = max(worst, float((np.abs(got - t) / (1e-4 * np.maximum(1.0, np.abs(t)))).max()))
        print(f"   worst error against float64 in units of the 1e-4 tolerance: {worst:.3f}")
        print(f"round {rnd} {Path(path).name}: {e0.elapsed_time(e1) / 40 / n * 1e3:.3f} us/window, max |diff| to first {(out - ref).abs().max().item():.2e}", flush=True)
