"""A/B of experimental builds of the wav2vec2 layer-0 kernels: python tools/ab_l0.py lib1.so [lib2.so ...]
Each library is loaded through ctypes directly (segma_w2v2_layer0 only) and timed alternately on 256 windows; outputs
are compared with the first library's."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from segma_b200 import synth  # noqa: E402

n, Cc, L, step = 256, 512, 64000, 63680
T0 = (L - 10) // 5 + 1
pcm = torch.from_numpy(synth.synth_audio(step * (n - 1) + L, 0)).cuda()
g = torch.Generator(device="cuda").manual_seed(0)
w = torch.randn((Cc, 10), device="cuda", generator=g) * 0.3
gamma = 1.0 + 0.1 * torch.randn(Cc, device="cuda", generator=g)
beta = 0.1 * torch.randn(Cc, device="cuda", generator=g)
ss = torch.empty((n, Cc, 2), device="cuda")
out = torch.empty((n, T0 + 1, Cc), dtype=torch.float16, device="cuda")
st = torch.cuda.current_stream().cuda_stream
libs = []
for path in sys.argv[1:]:
    lib = C.CDLL(path)
    lib.segma_w2v2_layer0.restype = C.c_int
    lib.segma_w2v2_layer0.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    libs.append((path, lib))
ref = None
for rnd in range(3):
    for path, lib in libs:
        call = lambda: lib.segma_w2v2_layer0(pcm.data_ptr(), pcm.numel(), n, L, step, w.data_ptr(), gamma.data_ptr(), beta.data_ptr(), Cc,  # noqa: E731
                                             ss.data_ptr(), out.data_ptr(), T0 + 1, st)
        for _ in range(3):
            assert call() == 0
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            call()
        e1.record()
        torch.cuda.synchronize()
        if ref is None:
            ref = out.clone()
        t = e0.elapsed_time(e1) / 20 / n * 1e3
        print(f"round {rnd} {Path(path).name}: {t:.3f} us/window, {(4 * L + 2 * T0 * Cc) / t / 1e3:.0f} GB/s algorithmic, "
              f"max |diff| to first {(out.float() - ref.float()).abs().max().item():.2e}", flush=True)
