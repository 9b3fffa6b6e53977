"""HBM throughput of the activation-quantization kernels at the shape the reference publishes for its own sm_100 cast
kernel (16384 x 16384 bf16 -> mxfp8: 5.4-5.7 TB/s, /root/reference docs/source/workflows/training.md:392-403), plus the
decode shapes.  GB/s = (bytes read + bytes written) / time, CUDA events over 20 launches on rotating buffers (> L2)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ao_b200  # noqa: E402,F401

ops = torch.ops.ao_b200
PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def bench(name, fn, inputs, out_bytes, in_bytes, iters=20):
    for a in inputs[:2]:
        fn(*a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(*inputs[i % len(inputs)])
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    gbs = (in_bytes + out_bytes) / us / 1e3
    print(f"  {name:34s} {us:9.1f} us  {gbs:8.1f} GB/s  {gbs / PEAK:.3f} of measured HBM peak", flush=True)


for (M, K) in ((16384, 16384), (4096, 14336), (32, 14336)):
    print(f"M={M} K={K}")
    xs = [torch.randn(M, K, device="cuda").to(torch.bfloat16) for _ in range(3 if M > 1000 else 8)]
    inb = M * K * 2
    bench("int8 per-token", lambda x: ops.int8_quantize_rowwise(x), [(x,) for x in xs], M * K + M * 4, inb)
    bench("e4m3 per-token", lambda x: ops.fp8_quantize_rowwise(x), [(x,) for x in xs], M * K + M * 4, inb)
    bench("mxfp8 rceil + blocked scales", lambda x: ops.mxfp8_quantize(x, True), [(x,) for x in xs], M * K + M * K // 32, inb)
    bench("nvfp4 + blocked scales", lambda x: ops.nvfp4_quantize(x, None, True), [(x,) for x in xs], M * K // 2 + M * K // 16, inb)
    if K <= 49152:
        w = torch.ones(K, device="cuda", dtype=torch.bfloat16)
        bench("rmsnorm -> e4m3 per-token (fused)", lambda x: ops.rmsnorm_quantize_rowwise(x, w, 1e-5, 1), [(x,) for x in xs], M * K + M * 4, inb)
        if M * K <= 4096 * 14336:
            ups = [torch.randn(M, K, device="cuda").to(torch.bfloat16) for _ in xs]
            bench("silu*up -> e4m3 per-token (fused)", lambda g, u: ops.silu_mul_quantize_rowwise(g, u, 1), list(zip(xs, ups)), M * K + M * 4, 2 * inb)
    del xs
    torch.cuda.empty_cache()
