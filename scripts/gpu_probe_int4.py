"""GPU bring-up probe for the int4 path (run under gpurun; prints a self-contained report).

Stages (each runs in its own subprocess with a timeout so a hung kernel cannot hang the call):
  pack   : our pack/unpack vs aten._convert_weight_to_int4pack (bit-exact), + one-hot layout dump on mismatch
  simple : CUDA-core kernel vs aten._weight_int4pack_mm vs fp32 dequant-matmul
  tc     : tcgen05 kernel, same comparisons, many shapes
  diag   : one-hot activations through the tc kernel (which k does each MMA slot read?)
  bench  : Llama-3-8B linear stack, ours vs aten, eager + CUDA graph
"""
import argparse
import math
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def load():
    torch.ops.load_library(os.path.join(ROOT, "ao_b200", "lib", "ao_b200_torch.so"))
    return torch.ops.ao_b200


def sqnr(ref, out):
    ref = ref.float()
    out = out.float()
    num = ref.norm()
    den = (ref - out).norm()
    if den == 0:
        return float("inf")
    return float(20 * torch.log10(num / den))


def make_weight(N, K, g, seed=0):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    q = torch.randint(0, 16, (N, K), device="cuda", generator=gen, dtype=torch.int32)
    s = (torch.rand(N, K // g, device="cuda", generator=gen) * 0.01 + 0.002).to(torch.bfloat16)
    z = ((torch.rand(N, K // g, device="cuda", generator=gen) - 0.5) * 0.02).to(torch.bfloat16)
    q_u8 = (q[:, ::2] << 4 | q[:, 1::2]).to(torch.uint8).contiguous()
    sz = torch.stack([s, z], dim=-1).transpose(0, 1).contiguous()  # [K/g, N, 2]
    return q, q_u8, s, z, sz


def dequant_ref(q, s, z, g):
    # bf16(fma(q-8, s, z)): compute exactly in fp64 then round once to bf16
    N, K = q.shape
    w = (q.double() - 8) * s.double().repeat_interleave(g, 1) + z.double().repeat_interleave(g, 1)
    return w.float().to(torch.bfloat16)  # fp64->fp32 is exact here (<=20 significant bits)


def stage_pack(ops):
    ok = True
    for (N, K) in [(8, 128), (64, 1024), (4096, 4096)]:
        q, q_u8, *_ = make_weight(N, K, 32, seed=N + K)
        ref = torch.ops.aten._convert_weight_to_int4pack(q_u8, 8)
        ours = ops.int4_pack_tile4d(q_u8, 8)
        same = torch.equal(ref, ours)
        back = ops.int4_unpack_tile4d(ref)
        rt = torch.equal(back, q_u8)
        print(f"[pack] N={N} K={K} shape={tuple(ref.shape)} pack_equal={same} unpack_roundtrip={rt}")
        ok &= same and rt
        if not same:
            nm = (ref != ours).sum().item()
            print(f"   mismatching words: {nm}/{ref.numel()}")
    for ikt in (2, 4):
        q, q_u8, *_ = make_weight(16, 256, 32, seed=ikt)
        ref = torch.ops.aten._convert_weight_to_int4pack(q_u8, ikt)
        ours = ops.int4_pack_tile4d(q_u8, ikt)
        print(f"[pack] inner_k_tiles={ikt} equal={torch.equal(ref, ours)}")
        ok &= torch.equal(ref, ours)
    if not ok:
        # dump the true layout with one-hot probes: where does q[n,k]=15 land?
        N, K = 16, 256
        print("[pack] one-hot layout dump (n, k) -> (n8, ko, lane, word, bit)")
        for n in (0, 1, 7, 8, 9):
            for k in list(range(0, 40)) + [64, 127, 128, 129, 255]:
                q = torch.zeros(N, K, dtype=torch.int32, device="cuda")
                q[n, k] = 15
                q_u8 = (q[:, ::2] << 4 | q[:, 1::2]).to(torch.uint8).contiguous()
                ref = torch.ops.aten._convert_weight_to_int4pack(q_u8, 8)
                nz = ref.nonzero()
                assert nz.shape[0] == 1
                idx = nz[0].tolist()
                val = ref[tuple(idx)].item() & 0xFFFFFFFF
                bit = int(math.log2(val // 15))
                print(f"   ({n},{k}) -> {idx} bit {bit}")
    return ok


def gemm_case(ops, M, N, K, g, impl, bias=False, seed=0, verbose=True):
    q, q_u8, s, z, sz = make_weight(N, K, g, seed=seed)
    qdata = torch.ops.aten._convert_weight_to_int4pack(q_u8, 8)
    gen = torch.Generator(device="cuda").manual_seed(seed + 1)
    x = torch.randn(M, K, device="cuda", generator=gen, dtype=torch.float32).to(torch.bfloat16)
    b = torch.randn(N, device="cuda", generator=gen).to(torch.bfloat16) if bias else None
    w_ref = dequant_ref(q, s, z, g)
    y32 = x.float() @ w_ref.float().t()
    if b is not None:
        y32 = y32 + b.float()
    y_aten = torch.ops.aten._weight_int4pack_mm(x, qdata, g, sz)
    if b is not None:
        y_aten = y_aten + b
    y = ops.int4_tilepacked_linear(x, qdata, g, sz, b, N, impl)
    torch.cuda.synchronize()
    s_ours = sqnr(y32, y)
    s_aten = sqnr(y32, y_aten)
    s_cross = sqnr(y_aten, y)
    finite = bool(torch.isfinite(y.float()).all())
    maxerr = float((y.float() - y32).abs().max())
    if verbose:
        print(f"[gemm impl={impl}] M={M:4d} N={N:5d} K={K:5d} g={g:3d} bias={int(bias)} "
              f"sqnr(ours,fp32)={s_ours:6.1f} sqnr(aten,fp32)={s_aten:6.1f} sqnr(ours,aten)={s_cross:6.1f} "
              f"maxerr={maxerr:.4g} finite={finite}")
    return s_ours, s_aten


def stage_dequant(ops):
    ok = True
    for g in (32, 64, 128, 256):
        q, q_u8, s, z, sz = make_weight(256, 1024, g, seed=g)
        qdata = torch.ops.aten._convert_weight_to_int4pack(q_u8, 8)
        w = ops.int4_dequant_tile4d(qdata, sz, g)
        ref = dequant_ref(q, s, z, g)
        eq = torch.equal(w, ref)
        print(f"[dequant] g={g} exact={eq} mismatches={(w != ref).sum().item()}")
        ok &= eq
    return ok


def stage_simple(ops):
    ok = True
    for (M, N, K, g) in [(1, 128, 1024, 32), (5, 256, 2048, 32), (16, 4096, 4096, 32), (8, 1024, 1024, 128), (3, 64, 1024, 256)]:
        so, sa = gemm_case(ops, M, N, K, g, impl=2, bias=(M == 5))
        ok &= so > 40
    return ok


def stage_tc(ops):
    ok = True
    cases = [
        (1, 128, 1024, 32), (16, 128, 1024, 32), (32, 128, 1024, 32), (32, 256, 2048, 32),
        (1, 4096, 4096, 32), (32, 4096, 4096, 32), (7, 1024, 4096, 32), (32, 14336, 4096, 32),
        (32, 4096, 14336, 32), (17, 4096, 4096, 64), (33, 4096, 4096, 128), (64, 4096, 4096, 256),
        (100, 1024, 2048, 32), (128, 4096, 4096, 32), (200, 1024, 4096, 32), (2, 136, 1024, 32),
    ]
    for (M, N, K, g) in cases:
        so, sa = gemm_case(ops, M, N, K, g, impl=1, bias=(M in (7, 33)))
        ok &= so > 40
    # determinism / semaphore restore: same call twice must be bit-identical
    q, q_u8, s, z, sz = make_weight(4096, 4096, 32, seed=3)
    qdata = torch.ops.aten._convert_weight_to_int4pack(q_u8, 8)
    x = torch.randn(32, 4096, device="cuda").to(torch.bfloat16)
    y1 = ops.int4_tilepacked_linear(x, qdata, 32, sz, None, 4096, 1)
    y2 = ops.int4_tilepacked_linear(x, qdata, 32, sz, None, 4096, 1)
    print("[tc] repeat bit-identical:", torch.equal(y1, y2))
    ok &= torch.equal(y1, y2)
    return ok


def stage_diag(ops):
    # one-hot x[0,k]=1: y[0,n] must equal W^[n,k] exactly.  Reports which k' it actually equals.
    N, K, g = 128, 1024, 32
    q, q_u8, s, z, sz = make_weight(N, K, g, seed=11)
    qdata = torch.ops.aten._convert_weight_to_int4pack(q_u8, 8)
    w_ref = dequant_ref(q, s, z, g)  # [N, K]
    bad = 0
    for k in list(range(0, 34)) + [63, 64, 65, 127, 128, 129, 255, 511, 512, 1023]:
        x = torch.zeros(16, K, device="cuda", dtype=torch.bfloat16)
        x[0, k] = 1.0
        x[3, k] = 2.0
        y = ops.int4_tilepacked_linear(x, qdata, g, sz, None, N, 1)
        torch.cuda.synchronize()
        col = y[0].float()
        exact = torch.equal(y[0], w_ref[:, k]) and torch.equal(y[3].float(), 2 * w_ref[:, k].float())
        if not exact:
            bad += 1
            # which k' matches?
            match = [kk for kk in range(K) if torch.equal(y[0], w_ref[:, kk])]
            nz_rows = (y.float().abs().sum(1) > 0).nonzero().flatten().tolist()
            print(f"[diag] k={k}: MISMATCH; matching k'={match[:8]} nonzero token rows={nz_rows[:8]} "
                  f"y[0,:4]={col[:4].tolist()} ref={w_ref[:4, k].float().tolist()}")
    print(f"[diag] one-hot mismatches: {bad}")
    return bad == 0


LLAMA8B = [("q", 4096, 4096), ("k", 1024, 4096), ("v", 1024, 4096), ("o", 4096, 4096),
           ("gate", 14336, 4096), ("up", 14336, 4096), ("down", 4096, 14336)]


def stage_bench(ops, layers=32, g=32):
    torch.manual_seed(0)
    dev = "cuda"
    weights = []
    total_bytes = 0
    for l in range(layers):
        lw = []
        for name, N, K in LLAMA8B:
            qdata = torch.randint(-2**31, 2**31 - 1, (N // 8, K // 128, 32, 4), device=dev, dtype=torch.int32)
            sz = (torch.rand(K // g, N, 2, device=dev) * 0.01).to(torch.bfloat16)
            lw.append((N, K, qdata, sz))
            total_bytes += qdata.numel() * 4 + sz.numel() * 2
        weights.append(lw)
    print(f"[bench] {layers} layers, packed bytes = {total_bytes/1e9:.3f} GB")
    for M in (1, 32):
        xs = {4096: torch.randn(M, 4096, device=dev).to(torch.bfloat16),
              14336: torch.randn(M, 14336, device=dev).to(torch.bfloat16)}

        def run_ours(impl=1):
            for lw in weights:
                for (N, K, qdata, sz) in lw:
                    ops.int4_tilepacked_linear(xs[K], qdata, g, sz, None, N, impl)

        def run_aten():
            for lw in weights:
                for (N, K, qdata, sz) in lw:
                    torch.ops.aten._weight_int4pack_mm(xs[K], qdata, g, sz)

        for name, fn in (("ours", run_ours), ("aten", run_aten)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 10
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            print(f"[bench] M={M:2d} {name:5s} eager : {ms:8.3f} ms/step  {M/ms*1e3:9.1f} tok/s  {total_bytes/ms/1e6:8.1f} GB/s")
            # CUDA graph
            try:
                gr = torch.cuda.CUDAGraph()
                st = torch.cuda.Stream()
                st.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(st):
                    fn()
                torch.cuda.current_stream().wait_stream(st)
                torch.cuda.synchronize()
                with torch.cuda.graph(gr):
                    fn()
                for _ in range(3):
                    gr.replay()
                torch.cuda.synchronize()
                e0.record()
                for _ in range(iters):
                    gr.replay()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / iters
                print(f"[bench] M={M:2d} {name:5s} graph : {ms:8.3f} ms/step  {M/ms*1e3:9.1f} tok/s  {total_bytes/ms/1e6:8.1f} GB/s")
            except Exception as ex:  # noqa
                print(f"[bench] M={M} {name} graph capture failed: {type(ex).__name__}: {ex}")
    # per-shape timing with rotating weights (distinct layers => no L2 reuse)
    for M in (1, 32):
        xs = {4096: torch.randn(M, 4096, device=dev).to(torch.bfloat16),
              14336: torch.randn(M, 14336, device=dev).to(torch.bfloat16)}
        for idx, (name, N, K) in enumerate(LLAMA8B):
            if name in ("v", "up"):
                continue
            byts = weights[0][idx][2].numel() * 4 + weights[0][idx][3].numel() * 2
            for who in ("ours", "aten"):
                def one(l):
                    Nn, Kk, qdata, sz = weights[l][idx]
                    if who == "ours":
                        ops.int4_tilepacked_linear(xs[Kk], qdata, g, sz, None, Nn, 1)
                    else:
                        torch.ops.aten._weight_int4pack_mm(xs[Kk], qdata, g, sz)
                for l in range(layers):
                    one(l)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for rep in range(3):
                    for l in range(layers):
                        one(l)
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) / (3 * layers) * 1e3
                print(f"[shape] M={M:2d} {name:5s} N={N:5d} K={K:5d} {who:5s}: {us:8.2f} us  {byts/us/1e3:8.1f} GB/s")
    return True


def stage_ncu(ops):
    # short run for ncu: a few Llama-shaped launches at M=32 and M=1
    g = 32
    for M in (32, 1):
        for name, N, K in LLAMA8B:
            qdata = torch.randint(-2**31, 2**31 - 1, (N // 8, K // 128, 32, 4), device="cuda", dtype=torch.int32)
            sz = (torch.rand(K // g, N, 2, device="cuda") * 0.01).to(torch.bfloat16)
            x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
            for _ in range(2):
                ops.int4_tilepacked_linear(x, qdata, g, sz, None, N, 1)
    torch.cuda.synchronize()
    return True


STAGES = {"pack": stage_pack, "dequant": stage_dequant, "simple": stage_simple, "tc": stage_tc,
          "diag": stage_diag, "bench": stage_bench, "ncu": stage_ncu}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", default=None)
    ap.add_argument("--stages", default="pack,dequant,simple,diag,tc,bench")
    ap.add_argument("--timeout", type=int, default=240)
    a = ap.parse_args()
    if a.stage:
        ops = load()
        ok = STAGES[a.stage](ops)
        torch.cuda.synchronize()
        print(f"[{a.stage}] RESULT {'OK' if ok else 'FAIL'}")
        sys.exit(0 if ok else 1)
    print(torch.cuda.get_device_name(0), torch.version.cuda)
    results = {}
    for st in a.stages.split(","):
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--stage", st],
                               timeout=a.timeout, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            out = r.stdout
            rc = r.returncode
        except subprocess.TimeoutExpired as e:
            out = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
            rc = "TIMEOUT"
        print(out[-12000:])
        print(f"== stage {st}: rc={rc} ({time.time()-t0:.1f}s)")
        results[st] = rc
    print("SUMMARY", results)
