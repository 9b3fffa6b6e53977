#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_int4_gpu.py -x -q -m gpu 2>&1 | tail -4
echo "=== sweep (auto grid)"; timeout 150 python -u scripts/gpu_prof_int4.py sweep 2>&1 | tail -13
echo "=== bench"; timeout 400 python bench.py --steps 20 --warmup 3 2> gpurun_out/bench_o.err | tee gpurun_out/bench_o.json; tail -3 gpurun_out/bench_o.err
