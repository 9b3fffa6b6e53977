#!/usr/bin/env python
"""bench.py — headline benchmark of the quantized-linear hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--bs 32] [--impl ours|reference]

Metric (BASELINE.json): tok/s of the Llama-3-8B int4 weight-only (tile_packed_to_4d, group_size=32)
linear stack.  One "step" = one pass of all 32x7 quantized linears over a batch of `bs` tokens per GPU
(decode: one token per sequence).  `value` is the whole-job tok/s with inputs resident in HBM, timed with
CUDA events over K CUDA-graph replays; `e2e` is the same pass driven from pinned HOST buffers (H2D of the
step's activations + D2H of its result inside the timed region) through the public API
(quantize_ -> nn.Linear.forward -> tensor-subclass dispatch -> torch.ops.ao_b200).  The same JSON line
carries the bs=1 measurement (`bs1`), the roofline of the dominant kernel and a CPU baseline.

Weights are synthetic random-init of the real shapes (no checkpoints offline); 4.36 GB of packed
weights per step >> the 126 MB L2, so no L2 flush is needed between iterations.
Multi-GPU (torchrun): batch sharding, one NCCL broadcast of the packed weights at setup, no
collective in the forward; value = N*bs / max-over-ranks time.

--impl reference: the reference's CPU implementation of the path, restated in oracle/ao_oracle.c
(the torchao Python package cannot travel to the GPU box), timed on all host cores on a
bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GROUP = 32


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _algo_bytes(shape, layers, bs):
    """0.625 B/param (int4 + (s,z) bf16 per 32) + activations in + out, per step."""
    w = shape.params_per_layer() * layers * (0.5 + 4.0 / GROUP)
    act = sum(bs * k * 2 + bs * n * 2 for _, n, k in shape.linears()) * layers
    return w + act


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import ao_b200  # noqa: F401  (loads the native library; raises if missing)
    from ao_b200.models import LLAMA3_8B, LlamaLinearStack
    from ao_b200.quantization import Int4WeightOnlyConfig, quantize_

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    shape = LLAMA3_8B
    layers = args.layers or shape.layers

    # ---- setup: build, quantize through the public API, replicate the packed weights ----------
    model = LlamaLinearStack(shape, layers=layers, device=dev, seed=0)
    quantize_(model, Int4WeightOnlyConfig(group_size=GROUP, int4_packing_format="tile_packed_to_4d"))
    torch.cuda.empty_cache()
    bcast_bytes = 0
    if world > 1:
        from ao_b200.parallel import broadcast_packed_weights

        bcast_bytes = broadcast_packed_weights(model, src=0)  # the one collective of the whole job
        torch.cuda.synchronize()

    launch_count = torch.ops.ao_b200.launch_count

    def measure(bs):
        gen = torch.Generator(device=dev).manual_seed(1 + rank)
        x_static = (torch.randn(bs, shape.hidden, device=dev, generator=gen)).to(torch.bfloat16)
        x_host = x_static.cpu().pin_memory()
        y_host = torch.empty(bs, shape.hidden, dtype=torch.bfloat16).pin_memory()
        # eager warm-up (also allocates the split-K workspace outside of capture)
        with torch.no_grad():
            for _ in range(2):
                model(x_static)
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            model(x_static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        c0 = launch_count()
        with torch.cuda.graph(graph), torch.no_grad():
            y_static = model(x_static)
        launches_per_step = launch_count() - c0

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        # ---- device-resident timing --------------------------------------------------------
        for _ in range(args.warmup):
            graph.replay()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        e0.record()
        for _ in range(args.steps):
            graph.replay()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if rank == 0 else None
        # ---- end-to-end: host buffers, H2D + D2H inside the timed region -----------------------
        for _ in range(args.warmup):
            x_static.copy_(x_host, non_blocking=True)
            graph.replay()
            y_host.copy_(y_static, non_blocking=True)
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        for _ in range(args.steps):
            x_static.copy_(x_host, non_blocking=True)
            graph.replay()
            y_host.copy_(y_static, non_blocking=True)
        e3.record()
        barrier()
        ms_e2e = e2.elapsed_time(e3)
        # ---- eager (no graph) end-to-end, for reference ----------------------------------------
        barrier()
        e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_eager = max(1, min(args.steps, 5))
        e4.record()
        with torch.no_grad():
            for _ in range(n_eager):
                y = model(x_host.to(dev, non_blocking=True))
                y_host.copy_(y, non_blocking=True)
        e5.record()
        barrier()
        ms_eager = e4.elapsed_time(e5) / n_eager
        t = torch.tensor([ms, ms_e2e, ms_eager], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, ms_eager = [float(v) for v in t.tolist()]
        finite = bool(torch.isfinite(y_static.float()).all())
        return {"ms_per_step": ms / args.steps, "ms_per_step_e2e": ms_e2e / args.steps, "ms_per_step_eager": ms_eager,
                "launches_per_step": int(launches_per_step), "clocks": clocks, "finite": finite,
                "h2d": x_host.numel() * 2, "d2h": y_host.numel() * 2}

    main = measure(args.bs)
    bs1 = measure(1) if args.bs != 1 else main
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = _peaks()
    toks = world * args.bs
    value = toks / (main["ms_per_step"] * 1e-3)
    ab = _algo_bytes(shape, layers, args.bs)
    kernel_launches = main["launches_per_step"]
    achieved = ab / (main["ms_per_step"] * 1e-3) / 1e9
    out = {
        "metric": "tok/s Llama-3-8B int4-wo (tile_packed_to_4d, g=32) linear stack, decode",
        "value": value, "unit": "tok/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int4 weights -> bf16 MMA (f32 accumulate)", "data": "synthetic (random-init weights of the real shapes)",
        "config": {"workload": f"Llama-3-8B int4-wo tile_packed_to_4d g=32, {layers} layers x 7 linears, bs={args.bs}/GPU decode",
                   "bs_per_gpu": args.bs, "layers": layers, "parallelism": f"batch-shard x{world} (replicated weights)",
                   "l2": "inputs larger than L2 (4.36 GB packed weights per step)", "timing": "CUDA events over CUDA-graph replays, max over ranks",
                   "weight_broadcast_bytes": bcast_bytes},
        "e2e": {"value": toks / (main["ms_per_step_e2e"] * 1e-3), "unit": "tok/s", "h2d_bytes_per_step": main["h2d"],
                "d2h_bytes_per_step": main["d2h"], "mode": "pinned host -> H2D -> CUDA-graph replay of model.forward -> D2H",
                "eager_no_graph_tok_s": toks / (main["ms_per_step_eager"] * 1e-3)},
        "gpu_launches": kernel_launches * args.steps,
        "clocks": main["clocks"],
        "bs1": {"value": world * 1 / (bs1["ms_per_step"] * 1e-3), "unit": "tok/s", "ms_per_step": bs1["ms_per_step"],
                "e2e_value": world * 1 / (bs1["ms_per_step_e2e"] * 1e-3),
                "roofline_frac": _algo_bytes(shape, layers, 1) / (bs1["ms_per_step"] * 1e-3) / 1e9 / peak},
        # traffic: dram__bytes_read+write per launch from the ncu --set full capture of this kernel on the seven
        # Llama-3-8B shapes at bs=32 (profiles/r01_int4_final_ncu.md): 139.3 MB per layer / 7 launches
        "roofline": {"bound": "hbm", "kernel": "ao::tsg::ts_gemm_kernel<ao::int4k::Int4Fmt, 32>", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": 19.90e6 if args.bs == 32 else None,
                     "algorithmic_bytes_per_launch": ab / max(1, kernel_launches), "peak_source": peak_src,
                     "algorithmic_bytes_per_step": ab, "launches_per_step": kernel_launches},
        "finite_outputs": main["finite"],
    }
    out["cpu_baseline"] = cpu_baseline(sample_layers=1, bs=1)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def _cpu_sample(bs, layers_sampled, threads):
    """Oracle int4 linear (oracle/ao_oracle.c::ao_oracle_int4_linear) over `layers_sampled` Llama-3-8B layers."""
    import numpy as np

    from ao_b200.models import LLAMA3_8B
    from oracle import oracle as o

    os.environ["OMP_NUM_THREADS"] = str(threads)
    o.lib().ao_oracle_set_threads(int(threads))   # torchrun exports OMP_NUM_THREADS=1 before we start
    rng = np.random.default_rng(0)
    shape = LLAMA3_8B
    mats = []
    for _, n, k in shape.linears():
        qd = rng.integers(-2**31, 2**31 - 1, size=(n // 8, k // 128, 32, 4), dtype=np.int64).astype(np.int32)
        sz = (rng.random((k // GROUP, n, 2), dtype=np.float32) * 0.01).astype(np.float32)
        mats.append((n, k, qd, o.f32_to_bf16(sz)))
    xs = {k: o.f32_to_bf16(rng.standard_normal((bs, k), dtype=np.float32)) for k in (shape.hidden, shape.inter)}
    o.lib()  # load
    t0 = time.perf_counter()
    for _ in range(layers_sampled):
        for n, k, qd, sz in mats:
            o.int4_linear(xs[k], qd, sz, GROUP)
    dt = time.perf_counter() - t0
    return dt / layers_sampled


class _AtenCpuInt4:
    """The reference's own CPU implementation of the path: torchao's CPU int4 route (Int4OpaqueTensor,
    torchao/prototype/quantization/int4/int4_opaque_tensor.py:197,414; BASELINE config[0]) is two PyTorch-core ops,
    aten._convert_weight_to_int4pack_for_cpu + aten._weight_int4pack_mm_for_cpu, and PyTorch is on the GPU box.
    One Llama-3-8B layer of packed weights is built once and reused; every step multiplies fresh activations."""

    def __init__(self, bs, threads):
        import torch

        from ao_b200.models import LLAMA3_8B

        gen = torch.Generator().manual_seed(0)
        self.torch = torch
        self.mats = []
        for _, n, k in LLAMA3_8B.linears():
            q = torch.randint(0, 16, (n, k), dtype=torch.int32, generator=gen)
            packed = torch.ops.aten._convert_weight_to_int4pack_for_cpu(q, 1)
            sz = (torch.rand(k // GROUP, n, 2, generator=gen) * 0.01).to(torch.bfloat16)
            self.mats.append((k, packed, sz))
        self.xs = {k: torch.randn(bs, k, generator=gen).to(torch.bfloat16) for k in (LLAMA3_8B.hidden, LLAMA3_8B.inter)}
        # Thread count: torchrun exports OMP_NUM_THREADS=1, and os.cpu_count() can exceed what the container may use
        # (CPU quota): oversubscribed, this kernel is >10x slower.  Time one pass per candidate and keep the fastest.
        best = (float("inf"), 1)
        t = max(1, int(threads))
        while t >= 1:
            torch.set_num_threads(t)
            self.layer_seconds()
            dt = min(self.layer_seconds() for _ in range(3))
            if dt < best[0]:
                best = (dt, t)
            t //= 2
        self.threads = best[1]
        torch.set_num_threads(self.threads)

    def layer_seconds(self, repeats=1):
        t0 = time.perf_counter()
        for _ in range(repeats):
            for k, packed, sz in self.mats:
                self.torch.ops.aten._weight_int4pack_mm_for_cpu(self.xs[k], packed, GROUP, sz)
        return (time.perf_counter() - t0) / repeats


def _cpu_runner(bs, threads):
    """(seconds-per-layer callable, kind, description): the PyTorch-core CPU kernel the reference calls when it exists,
    else the oracle port."""
    try:
        r = _AtenCpuInt4(bs, threads)
        r.layer_seconds()
        return r.layer_seconds, "reference", (f"aten._weight_int4pack_mm_for_cpu (PyTorch-core kernel the reference's CPU int4 "
                                               f"path calls, int4_opaque_tensor.py:414), {r.threads} threads (fastest of "
                                               f"{threads}, /2, /4, ... on this host)"), r.threads
    except Exception:  # op missing in this torch build: time the oracle port instead
        return (lambda repeats=1: _cpu_sample(bs, repeats, threads)), "port", "oracle/ao_oracle.c int4_linear", threads


def cpu_baseline(sample_layers=4, bs=1):
    threads = os.cpu_count() or 1
    try:
        from ao_b200.models import LLAMA3_8B

        run, kind, what, used = _cpu_runner(bs, threads)
        run(1)
        t_layer = run(sample_layers)
        return {"value": bs / (t_layer * LLAMA3_8B.layers), "unit": "tok/s", "cores": used, "kind": kind,
                "sample": f"{sample_layers} passes over one Llama-3-8B layer (7 int4 g=32 linears) at bs={bs}, {what}, time x32"}
    except Exception as ex:  # pragma: no cover
        return {"value": None, "unit": "tok/s", "cores": threads, "kind": "port", "sample": f"failed: {ex}"}


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    from ao_b200.models import LLAMA3_8B

    run, kind, what, used = _cpu_runner(args.bs, threads)
    # bounded sample: each "step" = 2 passes over the 7 linears of one layer at the configured batch, extrapolated
    # to the 32-layer stack (the weights of one layer, 136 MB packed, already exceed the CPU caches)
    warm = max(3, args.warmup)
    times = []
    for i in range(warm + args.steps):
        t = run(2)
        if i >= warm:
            times.append(t)
    t_layer = sum(times) / len(times)
    ms_step = t_layer * LLAMA3_8B.layers * 1e3
    value = args.bs / (ms_step * 1e-3)
    sample = (f"per step: 2 passes over one Llama-3-8B layer (7 int4 g=32 linears) at bs={args.bs}, {what}, time x32")
    out = {"impl": "reference", "metric": "tok/s Llama-3-8B int4-wo (tile_packed_to_4d, g=32) linear stack, decode",
           "value": value, "unit": "tok/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
           "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "int4 weights -> bf16 (CPU)", "data": "synthetic",
           "config": {"workload": f"Llama-3-8B int4-wo g=32, 32 layers x 7 linears, bs={args.bs} decode (CPU, sampled)"},
           "cpu_baseline": {"value": value, "unit": "tok/s", "cores": used, "kind": kind, "sample": sample},
           "e2e": {"value": value, "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--bs", type=int, default=32, help="tokens per GPU per step (decode batch)")
    ap.add_argument("--layers", type=int, default=0, help="override layer count (debug only; default = 32)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    a = ap.parse_args()
    if a.warmup < 3:
        a.warmup = 3
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
