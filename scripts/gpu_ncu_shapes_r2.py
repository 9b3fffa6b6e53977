"""Launch set for the round-2 ncu captures: the four launches of a fused Llama-3-8B layer (q|k|v, o, gate|up, down) at
bs = 32 and bs = 1 (decode kernel), then the gate|up projection at 512 and 4096 tokens (prefill kernel)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
torch.ops.load_library(os.path.join(ROOT, "ao_b200", "lib", "ao_b200_torch.so"))
ops = torch.ops.ao_b200
g = 32
SHAPES = [("qkv", 6144, 4096), ("o", 4096, 4096), ("gate_up", 28672, 4096), ("down", 4096, 14336)]
ws = {}
for name, N, K in SHAPES:
    qd = torch.randint(-2**31, 2**31 - 1, (N // 8, K // 128, 32, 4), device="cuda", dtype=torch.int32)
    sz = (torch.rand(K // g, N, 2, device="cuda") * 0.01).to(torch.bfloat16)
    ws[name] = (qd, sz)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "decode"):
    for M in (32, 1):
        for name, N, K in SHAPES:
            x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
            qd, sz = ws[name]
            ops.int4_tilepacked_linear(x, qd, g, sz, None, N, 1)
        torch.cuda.synchronize()
if which in ("all", "prefill"):
    for M in (512, 4096):
        for name in ("gate_up", "down"):
            _, N, K = [s for s in SHAPES if s[0] == name][0]
            x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
            qd, sz = ws[name]
            ops.int4_tilepacked_linear(x, qd, g, sz, None, N, 1)
        torch.cuda.synchronize()
