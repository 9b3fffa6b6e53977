#!/bin/bash
# memcheck + racecheck of one small invocation of each GEMM kernel (slow tools: tiny shapes only)
for tool in memcheck racecheck; do
  echo "=== compute-sanitizer $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python scripts/gpu_sanitize_workload.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard|done|Error|error" | head -12
done
