"""Experiment: one encoder layer's GEMMs on a limited persistent grid while the attention of another batch runs on
the SMs left over, against the same work run back to back on the whole chip.
  SEGMA_GEMM_MAX_CTAS=100 python tools/run_overlap.py [windows]"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from segma_b200 import ops  # noqa: E402

nw = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T, H, d, ffn = 1500, 12, 768, 3072
M = nw * T
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((M, d), device="cuda", generator=g).to(torch.float16)
hmid = torch.randn((M, ffn), device="cuda", generator=g).to(torch.float16)
res = torch.randn((M, d), device="cuda", generator=g)
w_qkv = (torch.randn((3 * d, d), device="cuda", generator=g) * 0.03).to(torch.float16)
w_o = (torch.randn((d, d), device="cuda", generator=g) * 0.03).to(torch.float16)
w_1 = (torch.randn((ffn, d), device="cuda", generator=g) * 0.03).to(torch.float16)
w_2 = (torch.randn((d, ffn), device="cuda", generator=g) * 0.03).to(torch.float16)
b_qkv = torch.randn(3 * d, device="cuda", generator=g)
b_o = torch.randn(d, device="cuda", generator=g)
b_1 = torch.randn(ffn, device="cuda", generator=g)
o_qkv = torch.empty((M, 3 * d), dtype=torch.float16, device="cuda")
o_mid = torch.empty((M, ffn), dtype=torch.float16, device="cuda")
o_res = torch.empty((M, d), dtype=torch.float32, device="cuda")
qkv2 = (torch.randn((M, 3 * d), device="cuda", generator=g) * 0.5).to(torch.float16)
o_att = torch.empty((M, d), dtype=torch.float16, device="cuda")


def gemms():
    ops.linear(x, w_qkv, b_qkv, out=o_qkv)
    ops.linear(x, w_o, b_o, add_src=res, out=o_res)
    ops.linear(x, w_1, b_1, gelu=True, out=o_mid)
    ops.linear(hmid, w_2, b_o, add_src=res, out=o_res)


def attn():
    ops.attention(qkv2, nw, T, H, out=o_att)


s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
reps = 6


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def serial():
    for _ in range(reps):
        gemms()
        attn()


def overlapped():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur)
    s2.wait_stream(cur)
    for _ in range(reps):
        with torch.cuda.stream(s1):
            gemms()
        with torch.cuda.stream(s2):
            attn()
    cur.wait_stream(s1)
    cur.wait_stream(s2)


print("limit", os.environ.get("SEGMA_GEMM_MAX_CTAS"), f"serial {timed(serial) / reps:.3f} ms/layer   overlapped {timed(overlapped) / reps:.3f} ms/layer")
