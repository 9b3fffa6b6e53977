#!/bin/bash
# lowp SS kernel grid heuristics at M = 32 on the small shapes (where the library kernel is ahead), fused quantizer GB/s
mkdir -p gpurun_out
echo "=== parity (fused quantizers)"; timeout 900 python -m pytest tests/test_lowp_gpu.py -q -x -k "fused or rmsnorm or silu or quantizers" 2>&1 | tail -3
echo "=== quant bw"; timeout 300 python -u scripts/gpu_quant_bw.py 2>&1 | head -13
for env in "" "AO_B200_TS_MIN_UNITS=8" "AO_B200_TS_MIN_UNITS=16" "AO_B200_TS_CTAS_PER_SM=1" "AO_B200_TS_CTAS_PER_SM=1 AO_B200_TS_MIN_UNITS=8" "AO_B200_TS_CTAS_PER_SM=1 AO_B200_TS_MIN_UNITS=2"; do
  echo "--- $env"; env $env timeout 600 python -u scripts/gpu_lowp_vs_library.py fp8 2>&1 | grep "M=32\|M= 4 qkv\|M= 1 qkv"
done
