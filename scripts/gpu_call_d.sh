#!/bin/bash
mkdir -p gpurun_out
for f in 0 1 2 3; do
  echo "=== AO_B200_TS_FLAGS=$f"
  AO_B200_TS_FLAGS=$f timeout 100 python scripts/gpu_timeline.py 1 2>&1 | tail -7
done
AO_B200_TS_FLAGS=3 timeout 100 python scripts/gpu_probe_int4.py --stage tc 2>&1 | grep -E "RESULT|sqnr\(ours,fp32\)= *[0-3]" | tail -3
