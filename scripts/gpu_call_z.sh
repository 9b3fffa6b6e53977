#!/bin/bash
echo "=== M=1 4096x4096"; timeout 100 python scripts/gpu_timeline.py 1 4096x4096 2>&1 | tail -4
echo "=== M=1 14336x4096"; timeout 100 python scripts/gpu_timeline.py 1 14336x4096 2>&1 | tail -4
echo "=== M=32 1024x4096"; timeout 100 python scripts/gpu_timeline.py 32 1024x4096 2>&1 | tail -4
echo "=== M=32 14336x4096"; timeout 100 python scripts/gpu_timeline.py 32 14336x4096 2>&1 | tail -4
