#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -u scripts/gpu_probe_int4.py --stages tc 2>&1 | grep -E "RESULT|FAIL|rror|identical" | tail -3
echo "=== sweep"; timeout 150 python -u scripts/gpu_prof_int4.py sweep 2>&1 | grep "M= 1"
echo "=== lowp pytest"; timeout 500 python -m pytest tests/test_lowp_gpu.py -q -m gpu --timeout 60 2>&1 | tail -15
echo "=== lowp bench"; timeout 200 python -u scripts/gpu_probe_lowp.py --stage bench 2>&1 | tail -30
