"""A few launches of the int4 linear for an ncu capture (big streaming shapes)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
torch.ops.load_library(os.path.join(ROOT, "ao_b200", "lib", "ao_b200_torch.so"))
ops = torch.ops.ao_b200
g = 32
for (M, N, K) in [(1, 14336, 8192), (32, 14336, 8192), (1, 28672, 4096), (1, 4096, 4096)]:
    qd = torch.randint(-2**31, 2**31 - 1, (N // 8, K // 128, 32, 4), device="cuda", dtype=torch.int32)
    sz = (torch.rand(K // g, N, 2, device="cuda") * 0.01).to(torch.bfloat16)
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    for _ in range(2):
        ops.int4_tilepacked_linear(x, qd, g, sz, None, N, 1)
    torch.cuda.synchronize()
