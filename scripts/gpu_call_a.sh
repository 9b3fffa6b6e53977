#!/bin/bash
# one GPU lease: v2 int4 validation + timing, first bench line, lowp tests, ncu capture
mkdir -p gpurun_out
timeout 240 python scripts/gpu_probe_int4.py --stages diag,tc --timeout 100 > gpurun_out/v2_probe.log 2>&1
tail -32 gpurun_out/v2_probe.log
timeout 240 python scripts/gpu_prof_int4.py sweep > gpurun_out/v2_sweep.log 2>&1; cat gpurun_out/v2_sweep.log
timeout 500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; cat gpurun_out/bench_a.json; tail -5 gpurun_out/bench_a.err
timeout 400 python -m pytest tests/test_lowp_gpu.py -m gpu -x -q 2>&1 | tail -8
timeout 240 ncu --set full --clock-control none --import-source on -k regex:ts_gemm -c 4 -o gpurun_out/prof_int4_r1b python scripts/gpu_prof_int4.py ncu > gpurun_out/ncu2.log 2>&1; tail -3 gpurun_out/ncu2.log
