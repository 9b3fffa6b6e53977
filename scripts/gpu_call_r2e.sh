#!/bin/bash
mkdir -p gpurun_out
echo "=== correctness"; timeout 900 python -m pytest tests/test_int4_gpu.py tests/test_fusion_gpu.py tests/test_lowp_gpu.py -q -x 2>&1 | grep -v "^DEBUG\|^$\|Warning\|warnings.warn\|return Variable\|Consider using\|return float" | tail -5
echo "=== layer"; timeout 300 python -u scripts/gpu_int4_layer.py one fused 2>&1 | tail -2
echo "--- FLAGS=1"; AO_B200_TS_FLAGS=1 timeout 200 python -u scripts/gpu_int4_layer.py one fused 2>&1 | tail -2
echo "--- PRODUCERS=1"; AO_B200_TS_PRODUCERS=1 timeout 200 python -u scripts/gpu_int4_layer.py one fused 2>&1 | tail -2
echo "=== shapes"; timeout 300 python -u scripts/gpu_int4_layer.py shapes 2>&1 | tail -14
echo "=== timeline"; for s in 28672x4096; do timeout 120 python -u scripts/gpu_timeline.py 1,32 $s 2>&1 | tail -26; done
