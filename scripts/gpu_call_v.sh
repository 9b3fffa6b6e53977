#!/bin/bash
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:ts_gemm -c 14 -o gpurun_out/r01_int4_final -f python scripts/gpu_ncu_bench_shapes.py > gpurun_out/ncu_final.log 2>&1
tail -3 gpurun_out/ncu_final.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r01_bench_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log | cut -c1-300; wc -l gpurun_out/r01_bench_launches.csv
