#!/bin/bash
for f in 1 2 3; do for c in 1 2; do
  echo "=== FLAGS=$f CTAS_PER_SM=$c"
  AO_B200_TS_FLAGS=$f AO_B200_TS_CTAS_PER_SM=$c timeout 100 python -u scripts/gpu_prof_int4.py sweep 2>&1 | grep -E "M= 1 N=14336|M=32 N=14336 K= 8192"
done; done
