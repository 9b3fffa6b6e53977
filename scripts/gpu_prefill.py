"""Prefill-shaped runs (M >= 512) of the quantized linears: TFLOP/s against the measured bf16 tensor peak.
  python scripts/gpu_prefill.py [int4|fp8|int8|all] [M,M,...] [NxK,NxK,...]
AO_B200_NO_PREFILL=1 routes M > 128 through the decode kernel's 128-token blocks (the A/B baseline)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ao_b200  # noqa: E402,F401

ops = torch.ops.ao_b200
PEAK = 1677.5
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
except Exception:
    pass


def time_fn(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def run(fmt, Ms, shapes):
    for M in Ms:
        for (N, K) in shapes:
            x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
            lib = None
            if fmt == "int4":
                qd = torch.randint(-2**31, 2**31 - 1, (N // 8, K // 128, 32, 4), device="cuda", dtype=torch.int32)
                sz = ((torch.rand(K // 32, N, 2, device="cuda") - 0.5) * 0.004).to(torch.bfloat16)
                fn = lambda: ops.int4_tilepacked_linear(x, qd, 32, sz, None, N, 1)
                lib = lambda: torch.ops.aten._weight_int4pack_mm(x, qd, 32, sz)
                mul = 1.0
            elif fmt in ("fp8", "int8"):
                w = torch.randn(N, K, device="cuda").to(torch.bfloat16)
                if fmt == "fp8":
                    wq, ws = ops.fp8_quantize_rowwise(w)
                    xq, xs = ops.fp8_quantize_rowwise(x)
                    fn = lambda: ops.fp8_rowwise_linear(xq, xs.reshape(-1), wq, ws.reshape(-1), None)
                    lib = lambda: torch._scaled_mm(xq, wq.t(), scale_a=xs.reshape(-1, 1), scale_b=ws.reshape(1, -1), out_dtype=torch.bfloat16, use_fast_accum=True)
                else:
                    wq, ws = ops.int8_quantize_rowwise(w)
                    xq, xs = ops.int8_quantize_rowwise(x)
                    fn = lambda: ops.int8_dyn_linear(xq, xs.reshape(-1), wq, ws.reshape(-1), None)
                    lib = lambda: torch._int_mm(xq, wq.t())
                mul = 2.0
            else:
                continue
            torch.cuda.synchronize()
            us = time_fn(fn)
            tf = 2.0 * M * N * K / us / 1e6
            extra = ""
            if lib is not None:
                lus = time_fn(lib)
                extra = f"  library {lus:8.1f} us ({2.0 * M * N * K / lus / 1e6:7.1f} TF)"
            xb = torch.randn(M, K, device="cuda").to(torch.bfloat16)
            wb = torch.randn(N, K, device="cuda").to(torch.bfloat16)
            bus = time_fn(lambda: torch.nn.functional.linear(xb, wb))
            print(f"  {fmt:5s} M={M:5d} N={N:5d} K={K:5d}: {us:9.1f} us  {tf:7.1f} TFLOP/s  {tf / (PEAK * mul):.3f} of {PEAK * mul:.0f}{extra}   bf16 cuBLAS {bus:8.1f} us", flush=True)
            del x


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "int4"
    Ms = tuple(int(v) for v in sys.argv[2].split(",")) if len(sys.argv) > 2 else (512, 4096)
    shapes = ([tuple(int(v) for v in s.split("x")) for s in sys.argv[3].split(",")] if len(sys.argv) > 3
              else [(6144, 4096), (4096, 4096), (28672, 4096), (4096, 14336)])
    for f in (["int4", "fp8", "int8"] if which == "all" else [which]):
        run(f, Ms, shapes)
