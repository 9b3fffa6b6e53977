#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/ -q -m gpu --timeout 90 2>&1 | tail -4
echo "=== sweep auto"; timeout 150 python -u scripts/gpu_prof_int4.py sweep 2>&1 | tail -12
echo "=== sweep CTAS=2"; AO_B200_TS_CTAS_PER_SM=2 timeout 150 python -u scripts/gpu_prof_int4.py sweep 2>&1 | tail -12
echo "=== sweep CTAS=1"; AO_B200_TS_CTAS_PER_SM=1 timeout 150 python -u scripts/gpu_prof_int4.py sweep 2>&1 | tail -12
echo "=== sweep MIN_UNITS=2"; AO_B200_TS_MIN_UNITS=2 timeout 150 python -u scripts/gpu_prof_int4.py sweep 2>&1 | grep -E "N= 4096 K= 4096|N= 1024"
echo "=== sweep MIN_UNITS=16"; AO_B200_TS_MIN_UNITS=16 timeout 150 python -u scripts/gpu_prof_int4.py sweep 2>&1 | grep -E "N= 4096 K= 4096|N= 1024|K= 2048"
