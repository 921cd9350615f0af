"""A/B of an environment switch on one box: each GEMM shape is timed in steady state with the variable set to 0 and
to 1, alternating several rounds, so box-to-box variation (about +-4 %) and governor drift cancel.
  python tools/ab_gemm.py SEGMA_GEMM_EPI_REGS [rows]"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from segma_b200 import ops  # noqa: E402

var = sys.argv[1]
M = int(sys.argv[2]) if len(sys.argv) > 2 else 192000
d, ffn = 768, 3072
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((M, d), device="cuda", generator=g).to(torch.float16)
w_qkv = (torch.randn((3 * d, d), device="cuda", generator=g) * 0.03).to(torch.float16)
w_1 = (torch.randn((ffn, d), device="cuda", generator=g) * 0.03).to(torch.float16)
b_qkv = torch.randn(3 * d, device="cuda", generator=g)
b_1 = torch.randn(ffn, device="cuda", generator=g)
o_qkv = torch.empty((M, 3 * d), dtype=torch.float16, device="cuda")
o_mid = torch.empty((M, ffn), dtype=torch.float16, device="cuda")
hmid = torch.randn((M, ffn), device="cuda", generator=g).to(torch.float16)
resid = torch.randn((M, d), device="cuda", generator=g)
w_o = (torch.randn((d, d), device="cuda", generator=g) * 0.03).to(torch.float16)
w_2 = (torch.randn((d, ffn), device="cuda", generator=g) * 0.03).to(torch.float16)
b_o = torch.randn(d, device="cuda", generator=g)
o_res = torch.empty((M, d), dtype=torch.float32, device="cuda")
cases = {
    "qkv  n2304 k768": lambda: ops.linear(x, w_qkv, b_qkv, out=o_qkv),
    "fc1  n3072 k768 gelu": lambda: ops.linear(x, w_1, b_1, gelu=True, out=o_mid),
    "out  n768 k768 +src f32": lambda: ops.linear(x, w_o, b_o, add_src=resid, out=o_res),
    "fc2  n768 k3072 +src f32": lambda: ops.linear(hmid, w_2, b_o, add_src=resid, out=o_res),
}


def run(fn, seconds):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    iters = max(10, int(seconds * 1e3 / (e0.elapsed_time(e1) / 5)))
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for name, fn in cases.items():
    res = {"0": [], "1": []}
    os.environ[var] = "0"
    run(fn, 0.5)  # settle the clock governor on this kernel
    for rnd in range(4):
        for val in ("0", "1"):
            os.environ[var] = val
            res[val].append(run(fn, 0.4))
    a, b = sum(res["0"]) / 4, sum(res["1"]) / 4
    print(f"{name:24s} {var}=0: {a * 1e3:7.1f} us   =1: {b * 1e3:7.1f} us   ratio {b / a:.4f}   rounds0 "
          + " ".join(f"{v * 1e3:.0f}" for v in res["0"]) + "  rounds1 " + " ".join(f"{v * 1e3:.0f}" for v in res["1"]))
