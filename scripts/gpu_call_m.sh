#!/bin/bash
for c in 1 2; do
  echo "=== CTAS_PER_SM=$c"; AO_B200_TS_CTAS_PER_SM=$c timeout 100 python scripts/gpu_timeline.py 1 2>&1 | tail -5
  echo "=== CTAS_PER_SM=$c FLAGS=3"; AO_B200_TS_FLAGS=3 AO_B200_TS_CTAS_PER_SM=$c timeout 100 python scripts/gpu_timeline.py 1 2>&1 | tail -5
done
echo "=== 4096x4096"; AO_B200_TS_CTAS_PER_SM=1 timeout 100 python scripts/gpu_timeline.py 1 4096x4096 2>&1 | tail -5
