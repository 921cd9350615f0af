"""Average board power and SM clock while one kernel type runs in a loop (about 1.5 s each):
python tools/power_by_kernel.py"""
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from segma_b200 import ops  # noqa: E402

nw, T, H, d, ffn = 128, 1500, 12, 768, 3072
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
gam = torch.randn(d, device="cuda", generator=g)
cases = {
    "qkv gemm": lambda: ops.linear(x, w_qkv, b_qkv, out=o_qkv),
    "fc1 gemm + gelu": lambda: ops.linear(x, w_1, b_1, gelu=True, out=o_mid),
    "fc2 gemm + residual": lambda: ops.linear(hmid, w_2, b_o, add_src=res, out=o_res),
    "out-proj gemm + residual": lambda: ops.linear(x, w_o, b_o, add_src=res, out=o_res),
    "attention": lambda: ops.attention(qkv2, nw, T, H, out=o_att),
    "layernorm": lambda: ops.layernorm(res, gam, gam, out_f16=x),
}


def sample(stop, out):
    while not stop.is_set():
        r = subprocess.run(["nvidia-smi", "--query-gpu=power.draw,clocks.sm", "--format=csv,noheader,nounits", "-i", "0"],
                           capture_output=True, text=True)
        try:
            p, c = r.stdout.strip().split(",")
            out.append((float(p), float(c)))
        except ValueError:
            pass
        time.sleep(0.05)


for name, fn in cases.items():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    iters = max(10, int(1500.0 / e0.elapsed_time(e1)))
    stop, samples = threading.Event(), []
    th = threading.Thread(target=sample, args=(stop, samples))
    th.start()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    stop.set()
    th.join()
    ms = e0.elapsed_time(e1) / iters
    tail = samples[len(samples) // 2:]  # second half: steady state
    pw = sum(s[0] for s in tail) / max(len(tail), 1)
    ck = sum(s[1] for s in tail) / max(len(tail), 1)
    print(f"{name:26s} {ms:8.3f} ms/launch  {pw:6.0f} W  {ck:6.0f} MHz  ({len(tail)} samples)")
