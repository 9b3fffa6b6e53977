#!/bin/bash
# round 2, call C: where does the streaming phase lose time?  flag experiments + ncu source-level profile; new parity tests
mkdir -p gpurun_out
echo "=== baseline"; timeout 300 python -u scripts/gpu_int4_layer.py one fused 2>&1 | tail -2
for f in 4 2 6; do echo "--- AO_B200_TS_FLAGS=$f"; AO_B200_TS_FLAGS=$f timeout 200 python -u scripts/gpu_int4_layer.py one fused 2>&1 | tail -2; done
echo "=== new tests"; timeout 900 python -m pytest tests/test_parity_holes_gpu.py tests/test_fusion_gpu.py -q -x 2>&1 | grep -v "^DEBUG\|^$" | tail -30
echo "=== ncu"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:ts_gemm -s 3 -c 1 -f -o gpurun_out/r02_gateup_m32 python scripts/gpu_ncu_one.py 32 28672 4096 5 2>&1 | tail -3
