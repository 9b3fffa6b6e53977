"""Per-shape timing of the int4 linear under CUDA graphs (no CPU launch overhead, no L2 reuse).

  python scripts/gpu_prof_int4.py sweep      # table for the current AO_B200_INT4_DBG / AO_B200_NO_PDL
  python scripts/gpu_prof_int4.py all        # runs sweep for DBG=0..3 and PDL on/off in subprocesses
  python scripts/gpu_prof_int4.py ncu        # a few launches for ncu
"""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def load():
    torch.ops.load_library(os.path.join(ROOT, "ao_b200", "lib", "ao_b200_torch.so"))
    return torch.ops.ao_b200


def time_graph(fn, iters=5):
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        fn()
    torch.cuda.current_stream().wait_stream(st)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def sweep(ops):
    g = 32
    copies = 24
    shapes = [(4096, 4096), (1024, 4096), (14336, 4096), (4096, 14336), (14336, 2048), (14336, 8192)]
    print(f"DBG={os.environ.get('AO_B200_INT4_DBG','0')} NO_PDL={os.environ.get('AO_B200_NO_PDL','0')}")
    for M in (1, 32):
        for (N, K) in shapes:
            ws = []
            for c in range(copies):
                qd = torch.randint(-2**31, 2**31 - 1, (N // 8, K // 128, 32, 4), device="cuda", dtype=torch.int32)
                sz = (torch.rand(K // g, N, 2, device="cuda") * 0.01).to(torch.bfloat16)
                ws.append((qd, sz))
            x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
            byts = ws[0][0].numel() * 4 + ws[0][1].numel() * 2

            def fn():
                for qd, sz in ws:
                    ops.int4_tilepacked_linear(x, qd, g, sz, None, N, 1)

            us = time_graph(fn) / copies
            print(f"  M={M:2d} N={N:5d} K={K:5d}: {us:8.2f} us/launch  {byts/us/1e3:8.1f} GB/s")
            del ws
            torch.cuda.empty_cache()


def ncu_stage(ops):
    g = 32
    for M in (1, 32):
        for (N, K) in [(14336, 4096), (4096, 4096), (4096, 14336)]:
            qd = torch.randint(-2**31, 2**31 - 1, (N // 8, K // 128, 32, 4), device="cuda", dtype=torch.int32)
            sz = (torch.rand(K // g, N, 2, device="cuda") * 0.01).to(torch.bfloat16)
            x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
            ops.int4_tilepacked_linear(x, qd, g, sz, None, N, 1)
    torch.cuda.synchronize()


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "sweep"
    if mode == "all":
        dbgs = sys.argv[2].split(",") if len(sys.argv) > 2 else ["0", "1", "2", "3"]
        for pdl in ("0",):
            for dbg in dbgs:
                env = dict(os.environ, AO_B200_INT4_DBG=dbg, AO_B200_NO_PDL=pdl)
                try:
                    r = subprocess.run([sys.executable, os.path.abspath(__file__), "sweep"], env=env, timeout=300,
                                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
                    print(r.stdout[-6000:])
                except subprocess.TimeoutExpired:
                    print(f"TIMEOUT dbg={dbg} pdl={pdl}")
    elif mode == "ncu":
        ncu_stage(load())
    else:
        sweep(load())
