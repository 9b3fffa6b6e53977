#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/ -q -m gpu --timeout 90 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "=== bench"; timeout 400 python bench.py 2> gpurun_out/bench_w.err | tee gpurun_out/bench_w.json | cut -c1-1500; tail -2 gpurun_out/bench_w.err
echo "=== reference arm"; timeout 400 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-700
echo "=== launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ts_gemm -c 448 --csv --log-file gpurun_out/r01_bench_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1; wc -l gpurun_out/r01_bench_launches.csv
