"""Per-shape decode GEMMs of the dynamic-activation formats against the library kernel the reference calls, M in {1, 4, 32}:
CUDA-graph chains of 16 different weights per shape (no L2 reuse), GEMM only (activations pre-quantized), us per launch.
  int8  lowp_linear_kernel<I8>   vs torch._int_mm (needs M > 16)
  fp8   lowp_linear_kernel<F8>   vs torch._scaled_mm rowwise
  mxfp8 lowp_linear_kernel<MXF8> vs torch._scaled_mm e8m0 block-scaled
  nvfp4 lowp_linear_kernel<NVF4> vs torch._scaled_mm fp4 block-scaled"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
torch.ops.load_library(os.path.join(ROOT, "ao_b200", "lib", "ao_b200_torch.so"))
ops = torch.ops.ao_b200
SHAPES = [("qkv", 6144, 4096), ("o", 4096, 4096), ("gate_up", 28672, 4096), ("down", 4096, 14336), ("k", 1024, 4096)]
COPIES = 16


def tg(fn, iters=5):
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        fn()
    torch.cuda.current_stream().wait_stream(st)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3 / COPIES


fmts = sys.argv[1].split(",") if len(sys.argv) > 1 else ["fp8", "int8", "mxfp8", "nvfp4"]
wins = {}
for fmt in fmts:
    for M in (1, 4, 32):
        for name, N, K in SHAPES:
            x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
            ws = [(torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16) for _ in range(2)]
            lib = None
            if fmt == "fp8":
                xq, xs = ops.fp8_quantize_rowwise(x)
                wq = [ops.fp8_quantize_rowwise(w) for w in ws]
                wq = [(q.clone(), s) for q, s in wq for _ in range(COPIES // 2)]
                ours = lambda: [ops.fp8_rowwise_linear(xq, xs.reshape(-1), q, s.reshape(-1), None) for q, s in wq]
                lib = lambda: [torch._scaled_mm(xq, q.t(), scale_a=xs.reshape(-1, 1), scale_b=s.reshape(1, -1), out_dtype=torch.bfloat16, use_fast_accum=True) for q, s in wq]
            elif fmt == "int8":
                xq, xs = ops.int8_quantize_rowwise(x)
                wq = [ops.int8_quantize_rowwise(w) for w in ws]
                wq = [(q.clone(), s) for q, s in wq for _ in range(COPIES // 2)]
                ours = lambda: [ops.int8_dyn_linear(xq, xs.reshape(-1), q, s.reshape(-1), None) for q, s in wq]
                if M > 16:
                    lib = lambda: [torch._int_mm(xq, q.t()) for q, s in wq]
            elif fmt == "mxfp8":
                xq, xs = ops.mxfp8_quantize(x, True)
                wq = [ops.mxfp8_quantize(w, True) for w in ws]
                wq = [(q.clone(), s) for q, s in wq for _ in range(COPIES // 2)]
                ours = lambda: [ops.mxfp8_linear(xq, xs, q, s, None) for q, s in wq]
                lib = lambda: [torch._scaled_mm(xq, q.t(), scale_a=xs.view(torch.float8_e8m0fnu), scale_b=s.view(torch.float8_e8m0fnu), out_dtype=torch.bfloat16) for q, s in wq]
            else:
                xq, xs = ops.nvfp4_quantize(x, None, True)
                wq = [ops.nvfp4_quantize(w, None, True) for w in ws]
                wq = [(q.clone(), s) for q, s in wq for _ in range(COPIES // 2)]
                ours = lambda: [ops.nvfp4_linear(xq, xs, None, q, s, None, None) for q, s in wq]
                lib = lambda: [torch._scaled_mm(xq.view(torch.float4_e2m1fn_x2), q.view(torch.float4_e2m1fn_x2).t(), scale_a=xs.view(torch.float8_e4m3fn),
                                                scale_b=s.view(torch.float8_e4m3fn), out_dtype=torch.bfloat16) for q, s in wq]
            try:
                t_o = tg(ours)
            except Exception as ex:
                print(f"  {fmt:5s} M={M:2d} {name:8s}: ours failed: {str(ex)[:120]}", flush=True)
                continue
            line = f"  {fmt:5s} M={M:2d} {name:8s} {N:5d}x{K:5d}: ours {t_o:7.2f} us"
            if lib is not None:
                try:
                    t_l = tg(lib)
                    line += f"   library {t_l:7.2f} us   ours/library {t_o / t_l:5.2f}  {'WIN' if t_o <= t_l else 'lose'}"
                    w = wins.setdefault((fmt, M), [0, 0])
                    w[0] += t_o <= t_l
                    w[1] += 1
                except Exception as ex:
                    line += f"   library failed: {str(ex)[:80]}"
            print(line, flush=True)
            del wq, ws
            torch.cuda.empty_cache()
print("wins (ours <= library) per format / M:", {f"{k[0]} M={k[1]}": f"{v[0]}/{v[1]}" for k, v in wins.items()})
