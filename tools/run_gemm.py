"""Times the four encoder-layer GEMM shapes alone: python tools/run_gemm.py [rows]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from segma_b200 import ops  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 48000
d, ffn = 768, 3072
g = torch.Generator(device="cuda").manual_seed(0)
x = (torch.randn((M, d), device="cuda", generator=g)).to(torch.float16)
hmid = (torch.randn((M, ffn), device="cuda", generator=g)).to(torch.float16)
res = torch.randn((M, d), device="cuda", generator=g)
cases = {
    "qkv  n2304 k768": lambda: ops.linear(x, w_qkv, b_qkv, out=o_qkv),
    "out  n768 k768 +src f32": lambda: ops.linear(x, w_o, b_o, add_src=res, out=o_res),
    "fc1  n3072 k768 gelu": lambda: ops.linear(x, w_1, b_1, gelu=True, out=o_mid),
    "fc2  n768 k3072 +src f32": lambda: ops.linear(hmid, w_2, b_o, add_src=res, out=o_res),
    "fc1  n3072 k768 (no gelu)": lambda: ops.linear(x, w_1, b_1, out=o_mid),
    "qkv  n2304 k768 (gelu)": lambda: ops.linear(x, w_qkv, b_qkv, gelu=True, out=o_qkv),
    # cuBLAS on the same shapes (plain GEMM, no epilogue), as the energy-efficiency yardstick under the power cap
    "qkv  cublas": lambda: torch.mm(x, w_qkv_t, out=o_qkv),
    "fc1  cublas": lambda: torch.mm(x, w_1_t, out=o_mid),
    "fc2  cublas": lambda: torch.mm(hmid, w_2_t, out=o_f16),
}
w_qkv = (torch.randn((3 * d, d), device="cuda", generator=g) * 0.03).to(torch.float16)
w_o = (torch.randn((d, d), device="cuda", generator=g) * 0.03).to(torch.float16)
w_1 = (torch.randn((ffn, d), device="cuda", generator=g) * 0.03).to(torch.float16)
w_2 = (torch.randn((d, ffn), device="cuda", generator=g) * 0.03).to(torch.float16)
w_qkv_t, w_1_t, w_2_t = w_qkv.t(), w_1.t(), w_2.t()
o_f16 = torch.empty((M, d), dtype=torch.float16, device="cuda")
b_qkv = torch.randn(3 * d, device="cuda", generator=g)
b_o = torch.randn(d, device="cuda", generator=g)
b_1 = torch.randn(ffn, device="cuda", generator=g)
o_qkv = torch.empty((M, 3 * d), dtype=torch.float16, device="cuda")
o_mid = torch.empty((M, ffn), dtype=torch.float16, device="cuda")
o_res = torch.empty((M, d), dtype=torch.float32, device="cuda")
flops = {"qkv": 2.0 * M * d * 3 * d, "out": 2.0 * M * d * d, "fc1": 2.0 * M * d * ffn, "fc2": 2.0 * M * d * ffn}
# The clock governor of a power-capped B200 moves in steps every 30-70 ms: each case runs long enough (about 0.4 s of
# warm-up, then 0.4 s timed) to be measured at its own steady-state clock rather than at its predecessor's.
for name, fn in cases.items():
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    iters = max(20, int(400.0 / (e0.elapsed_time(e1) / 10)))
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / iters * 1e-3
    print(f"{name:28s} {t * 1e6:8.1f} us  {flops[name[:3]] / t / 1e12:7.1f} TFLOP/s  ({iters} iterations)")
