#!/bin/bash
timeout 300 python -m pytest tests/test_int4_gpu.py tests/test_lowp_gpu.py -q -m gpu --timeout 120 2>&1 | tail -4
echo "=== lowp bench (2 CTAs/SM default)"; timeout 200 python -u scripts/gpu_probe_lowp.py --stage bench 2>&1 | grep -E "M=32|M= 1 gate|M= 1 down" | head -24
echo "=== lowp bench 1 CTA/SM"; AO_B200_TS_CTAS_PER_SM=1 timeout 200 python -u scripts/gpu_probe_lowp.py --stage bench 2>&1 | grep -E "M=32|M= 1 gate|M= 1 down" | grep -v torch | head -16
