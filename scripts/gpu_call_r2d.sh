#!/bin/bash
mkdir -p gpurun_out
echo "=== correctness"; timeout 600 python -m pytest tests/test_int4_gpu.py tests/test_fusion_gpu.py -q -x 2>&1 | grep -v "^DEBUG\|^$" | tail -4
echo "=== baseline (2 producers, 6 chunk slots)"; timeout 300 python -u scripts/gpu_int4_layer.py one fused 2>&1 | tail -2
echo "--- AO_B200_TS_PRODUCERS=1"; AO_B200_TS_PRODUCERS=1 timeout 200 python -u scripts/gpu_int4_layer.py one fused 2>&1 | tail -2
for f in 1 3 7; do echo "--- AO_B200_TS_FLAGS=$f"; AO_B200_TS_FLAGS=$f timeout 200 python -u scripts/gpu_int4_layer.py one fused 2>&1 | tail -2; done
echo "=== shapes"; timeout 300 python -u scripts/gpu_int4_layer.py shapes 2>&1 | tail -14
echo "=== timeline"; for s in 28672x4096 4096x14336 6144x4096; do timeout 120 python -u scripts/gpu_timeline.py 1,32 $s 2>&1 | tail -28; done
