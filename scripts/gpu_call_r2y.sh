#!/bin/bash
# nvfp4 scale tiles through a tensor-map TMA load: parity, config-5 shapes, compute-sanitizer
mkdir -p gpurun_out
echo "=== parity"; timeout 900 python -m pytest tests/test_lowp_gpu.py tests/test_fusion_gpu.py -q -x -k "nvfp4 or fused or Fused" 2>&1 | grep -E "passed|failed|FAILED"
echo "=== nvfp4-weight sweep"; timeout 300 python -u scripts/gpu_prof_nvfp4w.py 2>&1 | tail -5
echo "=== sanitizer"; bash scripts/gpu_sanitize.sh 2>&1 | tee gpurun_out/r02_compute_sanitizer.log
