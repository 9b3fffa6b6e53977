"""The seven Llama-3-8B linears of one layer at bs=32 and bs=1, once each after a warm-up, for ncu."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
torch.ops.load_library(os.path.join(ROOT, "ao_b200", "lib", "ao_b200_torch.so"))
ops = torch.ops.ao_b200
g = 32
SHAPES = [("q", 4096, 4096), ("k", 1024, 4096), ("v", 1024, 4096), ("o", 4096, 4096), ("gate", 14336, 4096),
          ("up", 14336, 4096), ("down", 4096, 14336)]
ws = {}
for name, N, K in SHAPES:
    qd = torch.randint(-2**31, 2**31 - 1, (N // 8, K // 128, 32, 4), device="cuda", dtype=torch.int32)
    sz = (torch.rand(K // g, N, 2, device="cuda") * 0.01).to(torch.bfloat16)
    ws[name] = (qd, sz)
for M in (32, 1):
    for name, N, K in SHAPES:
        x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        qd, sz = ws[name]
        ops.int4_tilepacked_linear(x, qd, g, sz, None, N, 1)
    torch.cuda.synchronize()
