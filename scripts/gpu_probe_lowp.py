"""GPU bring-up probe for int8 / fp8 / mxfp8 / nvfp4 (quantisers + GEMMs). Run under gpurun."""
import argparse
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def load():
    torch.ops.load_library(os.path.join(ROOT, "ao_b200", "lib", "ao_b200_torch.so"))
    return torch.ops.ao_b200


def sqnr(ref, out):
    ref, out = ref.double(), out.double()
    d = (ref - out).norm()
    return float("inf") if d == 0 else float(20 * torch.log10(ref.norm() / d))


def bits(t):
    return t.detach().cpu().contiguous().view(torch.int16).numpy().view(np.uint16)


def from_blocked_t(b, H, W):
    from oracle import oracle as o
    return torch.from_numpy(o.from_blocked(b.cpu().numpy().reshape(-1), H, W))


def stage_quant(ops):
    from oracle import oracle as o
    ok = True
    torch.manual_seed(0)
    for (M, K) in [(1, 4096), (5, 256), (32, 4096), (130, 512), (32, 14336)]:
        x = (torch.randn(M, K, device="cuda") * torch.logspace(-2, 2, M, device="cuda").unsqueeze(1)).to(torch.bfloat16)
        if M >= 5:
            x[2] = 0
        xb = bits(x)
        q, s = ops.int8_quantize_rowwise(x)
        qo, so = o.int8_quantize_rowwise(xb)
        e1 = np.array_equal(q.cpu().numpy(), qo) and np.array_equal(s.cpu().numpy().reshape(-1), so)
        q, s = ops.fp8_quantize_rowwise(x)
        qo, so = o.fp8_quantize_rowwise(xb)
        qn = q.view(torch.uint8).cpu().numpy()
        # all-zero rows: reference yields NaN (0/0); compare NaN-ness there, bytes elsewhere
        e2 = np.array_equal(qn, qo) and np.array_equal(s.cpu().numpy().reshape(-1), so)
        q, s = ops.mxfp8_quantize(x, False)
        qo, so = o.mxfp8_quantize(xb)
        e3 = np.array_equal(q.view(torch.uint8).cpu().numpy(), qo) and np.array_equal(s.cpu().numpy(), so)
        q2, s2 = ops.mxfp8_quantize(x, True)
        e3b = np.array_equal(s2.cpu().numpy().reshape(-1), o.to_blocked(so).reshape(-1)) and torch.equal(q2.view(torch.uint8), q.view(torch.uint8))
        q, s = ops.nvfp4_quantize(x, None, False)
        qo, so = o.nvfp4_quantize(xb, None)
        e4 = np.array_equal(q.cpu().numpy(), qo) and np.array_equal(s.cpu().numpy(), so)
        pts = (x.float().abs().max() / (448.0 * 6.0)).reshape(1)
        q, s = ops.nvfp4_quantize(x, pts, True)
        qo, so = o.nvfp4_quantize(xb, float(pts.item()))
        e5 = np.array_equal(q.cpu().numpy(), qo) and np.array_equal(s.cpu().numpy().reshape(-1), o.to_blocked(so).reshape(-1))
        print(f"[quant] M={M} K={K} int8={e1} fp8={e2} mxfp8={e3} mxfp8_swz={e3b} nvfp4_1lvl={e4} nvfp4_2lvl_swz={e5}")
        ok &= e1 and e2 and e3 and e3b and e4 and e5
    return ok


SHAPES = [(1, 128, 512), (16, 256, 1024), (32, 4096, 4096), (7, 1024, 4096), (32, 14336, 4096), (32, 4096, 14336),
          (64, 4096, 4096), (128, 1024, 2048), (200, 1024, 4096), (3, 136, 1024)]


def stage_int8(ops):
    ok = True
    torch.manual_seed(1)
    for (M, N, K) in SHAPES:
        xq = torch.randint(-128, 128, (M, K), device="cuda", dtype=torch.int8)
        wq = torch.randint(-128, 128, (N, K), device="cuda", dtype=torch.int8)
        acc_ref = (xq.double() @ wq.double().t()).to(torch.int64)
        acc = ops.int8_mm_i32(xq, wq)
        exact = torch.equal(acc.to(torch.int64), acc_ref)
        sx = torch.rand(M, 1, device="cuda") * 0.01 + 1e-3
        sw = torch.rand(N, device="cuda") * 0.01 + 1e-3
        b = torch.randn(N, device="cuda").to(torch.bfloat16) if M % 2 else None
        y = ops.int8_dyn_linear(xq, sx, wq, sw, b)
        t = (acc_ref.float() * sx).to(torch.bfloat16).float() * sw
        if b is not None:
            t = t + b.float()
        y_ref = t.to(torch.bfloat16)
        same = (y == y_ref).float().mean().item()
        print(f"[int8] M={M:4d} N={N:5d} K={K:5d} acc_exact={exact} y_equal_frac={same:.5f} sqnr={sqnr(y_ref, y):.1f}")
        ok &= exact and same > 0.999
    return ok


def stage_fp8(ops):
    ok = True
    torch.manual_seed(2)
    for (M, N, K) in SHAPES:
        if N % 16:
            continue
        x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
        xq, sx = ops.fp8_quantize_rowwise(x)
        wq, sw = ops.fp8_quantize_rowwise(w)
        b = torch.randn(N, device="cuda").to(torch.bfloat16) if M % 2 else None
        y = ops.fp8_rowwise_linear(xq, sx, wq, sw.reshape(-1), b)
        ref64 = (xq.double() @ wq.double().t()) * sx.double() * sw.double().reshape(1, -1)
        if b is not None:
            ref64 = ref64 + b.double()
        s_or = sqnr(ref64, y)
        try:
            y_t = torch._scaled_mm(xq, wq.t(), scale_a=sx, scale_b=sw.reshape(1, -1), bias=b, out_dtype=torch.bfloat16, use_fast_accum=True)
            s_t = sqnr(ref64, y_t)
            s_x = sqnr(y_t, y)
        except Exception as ex:
            s_t, s_x = float("nan"), float("nan")
            print("   _scaled_mm failed:", type(ex).__name__, str(ex)[:100])
        print(f"[fp8 ] M={M:4d} N={N:5d} K={K:5d} sqnr(ours,fp64)={s_or:6.1f} sqnr(torch,fp64)={s_t:6.1f} sqnr(ours,torch)={s_x:6.1f}")
        ok &= s_or > 45
    return ok


def stage_mxfp8(ops):
    from oracle import oracle as o
    ok = True
    torch.manual_seed(3)
    for (M, N, K) in SHAPES:
        x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
        xq, xs = ops.mxfp8_quantize(x, True)
        wq, ws = ops.mxfp8_quantize(w, True)
        b = torch.randn(N, device="cuda").to(torch.bfloat16) if M % 2 else None
        y = ops.mxfp8_linear(xq, xs, wq, ws, b)
        xs_p = from_blocked_t(xs, M, K // 32).cuda()
        ws_p = from_blocked_t(ws, N, K // 32).cuda()
        xd = xq.double() * torch.pow(2.0, xs_p.double() - 127).repeat_interleave(32, 1)
        wd = wq.double() * torch.pow(2.0, ws_p.double() - 127).repeat_interleave(32, 1)
        ref64 = xd @ wd.t()
        if b is not None:
            ref64 = ref64 + b.double()
        s_or = sqnr(ref64, y)
        try:
            y_t = torch._scaled_mm(xq, wq.t(), scale_a=xs.view(torch.float8_e8m0fnu), scale_b=ws.view(torch.float8_e8m0fnu), bias=b, out_dtype=torch.bfloat16)
            s_t, s_x = sqnr(ref64, y_t), sqnr(y_t, y)
        except Exception as ex:
            s_t, s_x = float("nan"), float("nan")
            print("   _scaled_mm failed:", type(ex).__name__, str(ex)[:120])
        print(f"[mxf8] M={M:4d} N={N:5d} K={K:5d} sqnr(ours,fp64)={s_or:6.1f} sqnr(torch,fp64)={s_t:6.1f} sqnr(ours,torch)={s_x:6.1f}")
        ok &= s_or > 45
    return ok


E2M1 = torch.tensor([0, 0.5, 1, 1.5, 2, 3, 4, 6, -0.0, -0.5, -1, -1.5, -2, -3, -4, -6], dtype=torch.float64)


def fp4_dequant(q, s_plain, pts):
    lut = E2M1.to(q.device)
    lo = lut[(q & 15).long()]
    hi = lut[(q >> 4).long()]
    v = torch.stack([lo, hi], dim=-1).reshape(q.shape[0], -1)
    sc = s_plain.view(torch.float8_e4m3fn).double().repeat_interleave(16, 1)
    return v * sc * (pts.double() if pts is not None else 1.0)


def stage_nvfp4(ops):
    ok = True
    torch.manual_seed(4)
    for (M, N, K) in SHAPES:
        if K % 256:
            continue
        x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
        pa = (x.float().abs().max() / (448.0 * 6.0)).reshape(1)
        pb = (w.float().abs().max() / (448.0 * 6.0)).reshape(1)
        xq, xs = ops.nvfp4_quantize(x, pa, True)
        wq, ws = ops.nvfp4_quantize(w, pb, True)
        b = torch.randn(N, device="cuda").to(torch.bfloat16) if M % 2 else None
        y = ops.nvfp4_linear(xq, xs, pa, wq, ws, pb, b)
        xs_p = from_blocked_t(xs, M, K // 16).cuda()
        ws_p = from_blocked_t(ws, N, K // 16).cuda()
        ref64 = fp4_dequant(xq, xs_p, pa) @ fp4_dequant(wq, ws_p, pb).t()
        if b is not None:
            ref64 = ref64 + b.double()
        s_or = sqnr(ref64, y)
        s_q = sqnr(x.double() @ w.double().t() + (b.double() if b is not None else 0), y)
        try:
            y_t = torch._scaled_mm(xq.view(torch.float4_e2m1fn_x2), wq.view(torch.float4_e2m1fn_x2).t(),
                                   scale_a=xs.view(torch.float8_e4m3fn), scale_b=ws.view(torch.float8_e4m3fn), out_dtype=torch.bfloat16)
            y_t = (y_t.float() * (pa * pb)).to(torch.bfloat16)
            if b is not None:
                y_t = y_t + b
            s_t, s_x = sqnr(ref64, y_t), sqnr(y_t, y)
        except Exception as ex:
            s_t, s_x = float("nan"), float("nan")
            print("   _scaled_mm failed:", type(ex).__name__, str(ex)[:120])
        print(f"[nvf4] M={M:4d} N={N:5d} K={K:5d} sqnr(ours,fp64)={s_or:6.1f} sqnr(torch,fp64)={s_t:6.1f} sqnr(ours,torch)={s_x:6.1f} sqnr_vs_bf16_linear={s_q:5.1f}")
        ok &= s_or > 45
    return ok


def stage_bench(ops):
    LL = [("q", 4096, 4096), ("k", 1024, 4096), ("gate", 14336, 4096), ("down", 4096, 14336)]
    copies = 16

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
        return e0.elapsed_time(e1) / iters * 1e3

    for M in (1, 32):
        for name, N, K in LL:
            ws8 = [torch.randint(-128, 128, (N, K), device="cuda", dtype=torch.int8) for _ in range(copies)]
            xq = torch.randint(-128, 128, (M, K), device="cuda", dtype=torch.int8)
            sx = torch.rand(M, 1, device="cuda")
            sw = torch.rand(N, device="cuda")
            us = tg(lambda: [ops.int8_dyn_linear(xq, sx, w, sw, None) for w in ws8]) / copies
            print(f"[bench int8 ] M={M:2d} {name:5s}: {us:8.2f} us  {N*K/us/1e3:8.1f} GB/s")
            wsf = [w.view(torch.float8_e4m3fn) for w in ws8]
            xf = xq.view(torch.float8_e4m3fn)
            us = tg(lambda: [ops.fp8_rowwise_linear(xf, sx, w, sw, None) for w in wsf]) / copies
            print(f"[bench fp8  ] M={M:2d} {name:5s}: {us:8.2f} us  {N*K/us/1e3:8.1f} GB/s")
            try:
                us = tg(lambda: [torch._scaled_mm(xf, w.t(), scale_a=sx, scale_b=sw.reshape(1, -1), out_dtype=torch.bfloat16, use_fast_accum=True) for w in wsf]) / copies
                print(f"[bench fp8-torch] M={M:2d} {name:5s}: {us:8.2f} us  {N*K/us/1e3:8.1f} GB/s")
            except Exception as ex:
                print("   torch fp8 bench failed", str(ex)[:100])
            del ws8, wsf
            torch.cuda.empty_cache()
    return True


STAGES = {"quant": stage_quant, "int8": stage_int8, "fp8": stage_fp8, "mxfp8": stage_mxfp8, "nvfp4": stage_nvfp4, "bench": stage_bench}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", default=None)
    ap.add_argument("--stages", default="quant,int8,fp8,mxfp8,nvfp4,bench")
    ap.add_argument("--timeout", type=int, default=200)
    a = ap.parse_args()
    if a.stage:
        ops = load()
        ok = STAGES[a.stage](ops)
        torch.cuda.synchronize()
        print(f"[{a.stage}] RESULT {'OK' if ok else 'FAIL'}")
        sys.exit(0 if ok else 1)
    results = {}
    for st in a.stages.split(","):
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--stage", st], timeout=a.timeout,
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            out, rc = r.stdout, r.returncode
        except subprocess.TimeoutExpired as e:
            out = e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
            rc = "TIMEOUT"
        print(out[-10000:])
        print(f"== stage {st}: rc={rc} ({time.time()-t0:.1f}s)")
        results[st] = rc
    print("SUMMARY", results)
