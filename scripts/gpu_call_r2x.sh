#!/bin/bash
# lowp kernel with the flag poller: parity, per-shape vs library (fp8), configs; then compute-sanitizer on the round-2 kernels
mkdir -p gpurun_out
echo "=== parity"; for rep in 1 2; do timeout 900 python -m pytest tests/test_lowp_gpu.py tests/test_fusion_gpu.py tests/test_parity_holes_gpu.py -q -x 2>&1 | grep -E "passed|failed|FAILED"; done
echo "=== fp8 vs library"; timeout 600 python -u scripts/gpu_lowp_vs_library.py fp8 2>&1 | tail -17
echo "=== sanitizer"; bash scripts/gpu_sanitize.sh 2>&1 | tee gpurun_out/r02_compute_sanitizer.log
