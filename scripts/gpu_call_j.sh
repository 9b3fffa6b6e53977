#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ts_gemm -c 8 -o gpurun_out/int4_v8 -f python scripts/gpu_ncu_int4.py > gpurun_out/ncu_v8.log 2>&1
tail -5 gpurun_out/ncu_v8.log; ls -la gpurun_out/*.ncu-rep
