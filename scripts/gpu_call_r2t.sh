#!/bin/bash
# early stage release with a real data dependence on the loads: failure rate of the test that exposed the race, perf
mkdir -p gpurun_out
for rep in 1 2 3 4 5 6 7 8; do
  timeout 300 python -m pytest tests/test_lowp_gpu.py -q -x -k "nvfp4_weight_linear" 2>&1 | grep -E "passed|failed|AssertionError: assert" | head -3
done
echo "=== int4 tests x3"; for rep in 1 2 3; do timeout 600 python -m pytest tests/test_int4_gpu.py tests/test_fusion_gpu.py -q -x 2>&1 | grep -E "passed|failed" ; done
echo "=== layer chain"; timeout 300 python -u scripts/gpu_int4_layer.py one fused 2>&1 | tail -2
echo "=== shapes"; timeout 300 python -u scripts/gpu_int4_layer.py shapes 2>&1 | tail -12
echo "=== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "passed|failed|FAILED" | tail -3
