#!/bin/bash
# round 2, call A: correctness of the owner-gather protocol + fusion, then layer-chain sweeps
mkdir -p gpurun_out
echo "=== hang probe"; timeout 120 python -u scripts/gpu_hang_probe.py 4096 4096 32 2>&1 | tail -6
timeout 120 python -u scripts/gpu_hang_probe.py 1024 4096 1 2>&1 | tail -6
echo "=== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "=== layer chain"; timeout 300 python -u scripts/gpu_int4_layer.py one fused 2>&1 | tail -4
timeout 300 python -u scripts/gpu_int4_layer.py one unfused 2>&1 | tail -4
echo "=== shapes"; timeout 300 python -u scripts/gpu_int4_layer.py shapes 2>&1 | tail -14
echo "=== sweep"; timeout 1500 python -u scripts/gpu_int4_layer.py sweep fused 2>&1 | tail -60
echo "=== timeline"; for s in 6144x4096 28672x4096 4096x14336; do timeout 120 python -u scripts/gpu_timeline.py 1,32 $s 2>&1 | tail -10; done
