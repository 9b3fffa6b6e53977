"""Llama-3-8B layer chain of int4 linears under a CUDA graph, straight through torch.ops (no Python dispatch):
   x -> qkv -> (q slice) -> o -> gate_up -> (gate slice) -> down -> next layer
fused (4 launches / layer, concatenated weights, strided slices) or unfused (7 launches / layer).
LAYERS distinct weight sets (>> L2) so a replay never finds weights in L2.

  python scripts/gpu_int4_layer.py one [fused|unfused] [M,...]     # one process, current environment
  python scripts/gpu_int4_layer.py sweep                           # subprocess per environment setting
  python scripts/gpu_int4_layer.py shapes                          # per-shape chains (24 distinct weights each)
"""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
G = 32
LAYERS = 8
H, I, KV = 4096, 14336, 1024


def load():
    torch.ops.load_library(os.path.join(ROOT, "ao_b200", "lib", "ao_b200_torch.so"))
    return torch.ops.ao_b200


def mk(N, K):
    qd = torch.randint(-2**31, 2**31 - 1, (N // 8, K // 128, 32, 4), device="cuda", dtype=torch.int32)
    sz = ((torch.rand(K // G, N, 2, device="cuda") - 0.5) * 0.004).to(torch.bfloat16)
    return qd, sz


def time_graph(fn, iters=10):
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        fn()
    torch.cuda.current_stream().wait_stream(st)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def layer_chain(ops, fused, M):
    if fused:
        shapes = [(H + 2 * KV, H), (H, H), (2 * I, H), (H, I)]
    else:
        shapes = [(H, H), (KV, H), (KV, H), (H, H), (I, H), (I, H), (H, I)]
    layers = [[mk(n, k) for n, k in shapes] for _ in range(LAYERS)]
    x0 = (torch.randn(M, H, device="cuda") * 0.5).to(torch.bfloat16)
    lin = lambda x, w: ops.int4_tilepacked_linear(x, w[0], G, w[1], None, w[0].shape[0] * 8, 1)

    def fn():
        x = x0
        for L in layers:
            if fused:
                qkv = lin(x, L[0])
                o = lin(qkv[:, :H], L[1])
                gu = lin(o, L[2])
                x = lin(gu[:, :I], L[3])
            else:
                q = lin(x, L[0])
                lin(x, L[1])
                lin(x, L[2])
                o = lin(q, L[3])
                g = lin(o, L[4])
                lin(o, L[5])
                x = lin(g, L[6])
        return x

    us = time_graph(fn) / LAYERS
    nbytes = sum(w[0].numel() * 4 + w[1].numel() * 2 for w in layers[0])
    return us, nbytes


def one(args):
    ops = load()
    mode = args[0] if args else "fused"
    Ms = [int(v) for v in args[1].split(",")] if len(args) > 1 else [1, 32]
    for M in Ms:
        us, nbytes = layer_chain(ops, mode == "fused", M)
        print(f"  {mode:8s} M={M:3d}: {us:8.2f} us/layer  {nbytes/us/1e3:8.1f} GB/s  -> {us*32/1e3:6.3f} ms/step  frac {nbytes/us/1e3/6587.7:.3f}", flush=True)


def shapes(args):
    ops = load()
    copies = 24
    table = [(H + 2 * KV, H), (H, H), (2 * I, H), (H, I), (KV, H), (I, H)]
    for M in (1, 32):
        for (N, K) in table:
            ws = [mk(N, K) for _ in range(copies)]
            x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
            byts = ws[0][0].numel() * 4 + ws[0][1].numel() * 2

            def fn():
                for qd, sz in ws:
                    ops.int4_tilepacked_linear(x, qd, G, sz, None, N, 1)

            us = time_graph(fn, 5) / copies
            print(f"  M={M:2d} N={N:5d} K={K:5d}: {us:8.2f} us/launch  {byts/us/1e3:8.1f} GB/s", flush=True)
            del ws
            torch.cuda.empty_cache()


def sweep(args):
    settings = [
        {},
        {"AO_B200_TS_PREFETCH": "0"},
        {"AO_B200_TS_PREFETCH": "8"},
        {"AO_B200_TS_PREFETCH": "128"},
        {"AO_B200_TS_PREFETCH": "4096"},
        {"AO_B200_TS_CTAS_PER_SM": "1"},
        {"AO_B200_TS_CTAS_PER_SM": "2"},
        {"AO_B200_TS_MIN_UNITS": "8"},
        {"AO_B200_TS_MIN_UNITS": "2"},
        {"AO_B200_NO_PDL": "1"},
    ]
    modes = args if args else ["fused"]
    for env_add in settings:
        for mode in modes:
            print(f"=== {env_add} {mode}", flush=True)
            env = dict(os.environ, **env_add)
            try:
                r = subprocess.run([sys.executable, "-u", os.path.abspath(__file__), "one", mode], env=env, timeout=150,
                                   stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
                print(r.stdout[-1500:], flush=True)
            except subprocess.TimeoutExpired as e:
                print("TIMEOUT", (e.stdout or b"")[-500:], flush=True)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "one"
    {"one": one, "sweep": sweep, "shapes": shapes}[mode](sys.argv[2:])
