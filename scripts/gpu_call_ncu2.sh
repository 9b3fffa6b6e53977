#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"lowp_linear|ts_gemm|quant" -c 40 -o gpurun_out/r02_lowp_quant -f python scripts/gpu_ncu_lowp_r2.py > gpurun_out/ncu_lowp.log 2>&1; tail -2 gpurun_out/ncu_lowp.log; ls -la gpurun_out/r02_lowp_quant.ncu-rep
