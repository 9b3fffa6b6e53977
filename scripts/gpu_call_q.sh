#!/bin/bash
for m in 1 32; do
  echo "=== M=$m 1024x4096"; timeout 100 python scripts/gpu_timeline.py $m 1024x4096 2>&1 | tail -4
  echo "=== M=$m 4096x4096"; timeout 100 python scripts/gpu_timeline.py $m 4096x4096 2>&1 | tail -4
done
