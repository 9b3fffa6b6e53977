#!/bin/bash
# single-poller flag prefetch (activation-producer warp polls, cta-scope barrier for the others): parity + perf
mkdir -p gpurun_out
echo "=== int4 tests x2"; for rep in 1 2; do timeout 600 python -m pytest tests/test_int4_gpu.py tests/test_fusion_gpu.py tests/test_lowp_gpu.py -q -x 2>&1 | grep -E "passed|failed|FAILED" ; done
echo "=== layer chain"; timeout 300 python -u scripts/gpu_int4_layer.py one fused 2>&1 | tail -2
echo "=== shapes"; timeout 300 python -u scripts/gpu_int4_layer.py shapes 2>&1 | tail -12
echo "=== hang probe"; timeout 300 python -u scripts/gpu_hang_probe.py 2>&1 | tail -6
