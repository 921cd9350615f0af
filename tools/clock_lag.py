import sys
from pathlib import Path
import torch
sys.path.insert(0, "/root/repo")
from segma_b200 import ops
nw, T, H, d, ffn = 128, 1500, 12, 768, 3072
M = nw * T
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((M, d), device="cuda", generator=g).to(torch.float16)
hmid = torch.randn((M, ffn), device="cuda", generator=g).to(torch.float16)
res = torch.randn((M, d), device="cuda", generator=g)
w_qkv = (torch.randn((3 * d, d), device="cuda", generator=g) * 0.03).to(torch.float16)
w_1 = (torch.randn((ffn, d), device="cuda", generator=g) * 0.03).to(torch.float16)
w_2 = (torch.randn((d, ffn), device="cuda", generator=g) * 0.03).to(torch.float16)
b_qkv = torch.randn(3 * d, device="cuda", generator=g); b_o = torch.randn(d, device="cuda", generator=g); b_1 = torch.randn(ffn, device="cuda", generator=g)
o_qkv = torch.empty((M, 3 * d), dtype=torch.float16, device="cuda")
o_mid = torch.empty((M, ffn), dtype=torch.float16, device="cuda")
o_res = torch.empty((M, d), dtype=torch.float32, device="cuda")
qkv2 = (torch.randn((M, 3 * d), device="cuda", generator=g) * 0.5).to(torch.float16)
o_att = torch.empty((M, d), dtype=torch.float16, device="cuda")
def gemms():
    ops.linear(x, w_qkv, b_qkv, out=o_qkv); ops.linear(x, w_1, b_1, gelu=True, out=o_mid); ops.linear(hmid, w_2, b_o, add_src=res, out=o_res)
def attn(): ops.attention(qkv2, nw, T, H, out=o_att)
for _ in range(3): gemms(); attn()
torch.cuda.synchronize()
# phase 1: 60 gemm groups (~140 ms), then 40 attention launches timed individually, then 40 gemm groups timed
ev = []
for _ in range(60): gemms()
for i in range(40):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); attn(); e1.record(); ev.append(("attn", e0, e1))
for i in range(40):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gemms(); e1.record(); ev.append(("gemm", e0, e1))
# interleaved like the pipeline
for i in range(20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gemms(); e1.record(); ev.append(("gemm_i", e0, e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); attn(); e1.record(); ev.append(("attn_i", e0, e1))
torch.cuda.synchronize()
for kind in ("attn", "gemm", "attn_i", "gemm_i"):
    ts = [e0.elapsed_time(e1) for k, e0, e1 in ev if k == kind]
    print(kind, " ".join(f"{t:.2f}" for t in ts))
