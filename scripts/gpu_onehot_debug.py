"""One-hot exactness of the int4 linear, with a report of WHERE it fails:  python scripts/gpu_onehot_debug.py M N K"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
torch.ops.load_library(os.path.join(ROOT, "ao_b200", "lib", "ao_b200_torch.so"))
ops = torch.ops.ao_b200
M, N, K = (int(v) for v in sys.argv[1:4])
g = 32
gen = torch.Generator(device="cuda").manual_seed(11)
q = torch.randint(0, 16, (N, K), device="cuda", generator=gen, dtype=torch.int32)
s = (torch.rand(N, K // g, device="cuda", generator=gen) * 0.01 + 0.002).to(torch.bfloat16)
z = ((torch.rand(N, K // g, device="cuda", generator=gen) - 0.5) * 0.02).to(torch.bfloat16)
q_u8 = (q[:, ::2] << 4 | q[:, 1::2]).to(torch.uint8).contiguous()
sz = torch.stack([s, z], dim=-1).transpose(0, 1).contiguous()
qd = ops.int4_pack_tile4d(q_u8, 8)
w = ops.int4_dequant_tile4d(qd, sz, g)
x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
for rep in range(3):
    y = ops.int4_tilepacked_linear(x, qd, g, sz, None, N, 0)    # a random launch in between, like the test
    rows = sorted({0, 255 % M, 256 % M, M - 1})
    ks = [(977 * (i + 1)) % K for i in range(len(rows))]
    xh = torch.zeros(M, K, device="cuda", dtype=torch.bfloat16)
    for m, k in zip(rows, ks):
        xh[m, k] = 1.0
    yh = ops.int4_tilepacked_linear(xh, qd, g, sz, None, N, 0)
    torch.cuda.synchronize()
    expect = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    for m, k in zip(rows, ks):
        expect[m] = w[:, k]
    bad = (yh != expect)
    print(f"rep {rep}: {int(bad.sum())} mismatching outputs of {M * N}")
    if bad.any():
        idx = bad.nonzero()
        ms, ns = idx[:, 0], idx[:, 1]
        print("   rows:", sorted(set(ms.tolist()))[:20], " n tiles:", sorted(set((ns // 128).tolist()))[:40])
        for i in range(min(6, idx.shape[0])):
            m, n = int(ms[i]), int(ns[i])
            print(f"   y[{m},{n}] = {float(yh[m, n]):.6g}  expected {float(expect[m, n]):.6g}   y_random[{m},{n}] = {float(y[m, n]):.6g}")
