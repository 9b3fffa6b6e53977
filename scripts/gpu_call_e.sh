#!/bin/bash
# v5 (multi-issuer) bring-up: correctness first, then sweep under each issuer mode
mkdir -p gpurun_out
timeout 200 python scripts/gpu_probe_int4.py --stage tc 2>&1 | grep -E "RESULT|FAIL|Error|error|sqnr\(ours,fp32\)= *[0-3]" | tail -5
for f in 0 4 8; do
  echo "=== AO_B200_TS_FLAGS=$f"
  AO_B200_TS_FLAGS=$f timeout 150 python scripts/gpu_prof_int4.py sweep 2>&1 | tail -14
done
AO_B200_TS_FLAGS=4 timeout 200 python scripts/gpu_probe_int4.py --stage tc 2>&1 | grep -E "RESULT|FAIL|Error|error" | tail -3
