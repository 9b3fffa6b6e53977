"""Per-shape timing of the nvfp4-weight linear (BASELINE config 5: nvfp4 weights x fp8 rowwise activations) on the
Llama-3-70B shapes, under CUDA graphs (chain of 8 different weights per shape)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scripts.gpu_prof_int4 import load, time_graph

ops = load()
M = 32
for name, N, K in [("q/o", 8192, 8192), ("k/v", 1024, 8192), ("gate/up", 28672, 8192), ("down", 8192, 28672)]:
    copies = 8
    ws = []
    for c in range(copies):
        w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
        pts = (w.float().abs().max() / (448.0 * 6.0)).reshape(1)
        wq, wsf = ops.nvfp4_quantize(w, pts, True)
        ws.append((wq, wsf, pts))
        del w
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    xq, sx = ops.fp8_fakequant_rowwise(x)
    sx = sx.reshape(-1)
    byts = ws[0][0].numel() + ws[0][1].numel()

    def fn():
        for wq, wsf, pts in ws:
            ops.nvfp4_weight_linear(xq, sx, wq, wsf, pts, None)

    us = time_graph(fn) / copies
    print(f"  nvfp4-w x fp8-act M={M} {name:8s} N={N:5d} K={K:5d}: {us:8.2f} us/launch  {byts/us/1e3:8.1f} GB/s")
    del ws
    torch.cuda.empty_cache()
