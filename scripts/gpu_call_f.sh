#!/bin/bash
mkdir -p gpurun_out
echo "=== mode A (flags 0) chain probe"; AO_B200_TS_FLAGS=0 timeout 60 python -u scripts/gpu_hang_probe.py 2>&1 | tail -12
echo "=== mode A, no PDL";            AO_B200_NO_PDL=1 AO_B200_TS_FLAGS=0 timeout 60 python -u scripts/gpu_hang_probe.py 2>&1 | tail -12
echo "=== mode A, 14336x4096";        AO_B200_TS_FLAGS=0 timeout 60 python -u scripts/gpu_hang_probe.py 14336 4096 1 2>&1 | tail -12
echo "=== mode 8 chain probe";        AO_B200_TS_FLAGS=8 timeout 60 python -u scripts/gpu_hang_probe.py 2>&1 | tail -12
echo "=== timeline flags 8"; AO_B200_TS_FLAGS=8 timeout 100 python scripts/gpu_timeline.py 1 2>&1 | tail -14
echo "=== timeline flags 0"; AO_B200_TS_FLAGS=0 timeout 100 python scripts/gpu_timeline.py 1 2>&1 | tail -14
echo "=== timeline flags 4"; AO_B200_TS_FLAGS=4 timeout 100 python scripts/gpu_timeline.py 1 2>&1 | tail -14
