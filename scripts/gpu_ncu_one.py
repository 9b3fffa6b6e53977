"""A few launches of one int4 linear shape for an ncu capture:  python scripts/gpu_ncu_one.py M N K [count]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
torch.ops.load_library(os.path.join(ROOT, "ao_b200", "lib", "ao_b200_torch.so"))
ops = torch.ops.ao_b200
M, N, K = (int(v) for v in sys.argv[1:4])
count = int(sys.argv[4]) if len(sys.argv) > 4 else 6
g = 32
ws = []
for c in range(count):
    qd = torch.randint(-2**31, 2**31 - 1, (N // 8, K // 128, 32, 4), device="cuda", dtype=torch.int32)
    sz = ((torch.rand(K // g, N, 2, device="cuda") - 0.5) * 0.004).to(torch.bfloat16)
    ws.append((qd, sz))
x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
for qd, sz in ws:
    ops.int4_tilepacked_linear(x, qd, g, sz, None, N, 1)
torch.cuda.synchronize()
