#!/bin/bash
# nvfp4-weight dequant: prmt extraction + one multiply (2^60 folded into the epilogue): parity and the 70B shapes
mkdir -p gpurun_out
echo "=== parity"; for rep in 1 2; do timeout 900 python -m pytest tests/test_lowp_gpu.py tests/test_fusion_gpu.py tests/test_parity_holes_gpu.py -q -x 2>&1 | grep -E "passed|failed|FAILED|assert"; done
echo "=== nvfp4-weight sweep (before: q/o 22.6, k/v 10.6, gate/up 58.4, down 57.5 us)"; timeout 300 python -u scripts/gpu_prof_nvfp4w.py 2>&1 | tail -5
