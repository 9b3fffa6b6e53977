#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -u scripts/gpu_probe_int4.py --stages diag,tc 2>&1 | grep -E "RESULT|FAIL|rror|identical|== stage|sqnr\(ours,fp32\)= *(-|nan|[0-3][0-9]\.)" | tail -8
for c in 2 1; do
  echo "=== AO_B200_TS_CTAS_PER_SM=$c"
  AO_B200_TS_CTAS_PER_SM=$c timeout 150 python -u scripts/gpu_prof_int4.py sweep 2>&1 | tail -13
done
