#!/bin/bash
mkdir -p gpurun_out
echo "=== nvfp4-weight sweep"; timeout 200 python -u scripts/gpu_prof_nvfp4w.py 2>&1 | tail -5
echo "=== ncu full"; timeout 500 ncu --set full --clock-control none --import-source on -k regex:ts_gemm -c 14 -o gpurun_out/r01_int4_final -f python scripts/gpu_ncu_bench_shapes.py > gpurun_out/ncu_final.log 2>&1; tail -2 gpurun_out/ncu_final.log
echo "=== bench"; timeout 400 python bench.py 2> gpurun_out/bench_final.err | tee gpurun_out/bench_final.json | cut -c1-400; tail -2 gpurun_out/bench_final.err
echo "=== launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ts_gemm -c 448 --csv --log-file gpurun_out/r01_bench_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1; wc -l gpurun_out/r01_bench_launches.csv
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
