#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -u scripts/gpu_probe_int4.py --stages diag,tc 2>&1 | grep -E "RESULT|FAIL|rror|identical" | tail -4
timeout 200 python -m pytest tests/test_int4_gpu.py -q -m gpu -x --timeout 60 2>&1 | tail -2
echo "=== sweep"; timeout 150 python -u scripts/gpu_prof_int4.py sweep 2>&1 | tail -12
echo "=== lowp bench"; timeout 200 python -u scripts/gpu_probe_lowp.py --stage bench 2>&1 | grep -E "M=32|gate|down" | head -24
