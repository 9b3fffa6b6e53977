#!/bin/bash
mkdir -p gpurun_out
echo "=== chain probe M=1"; timeout 60 python -u scripts/gpu_hang_probe.py 14336 4096 1 2>&1 | tail -4
echo "=== chain probe M=32"; timeout 60 python -u scripts/gpu_hang_probe.py 14336 4096 32 2>&1 | tail -4
timeout 200 python -u scripts/gpu_probe_int4.py --stage tc 2>&1 | grep -E "RESULT|FAIL|rror|identical|sqnr\(ours,fp32\)= *(-|nan|[0-3][0-9]\.)" | tail -8
echo "=== default"; timeout 150 python -u scripts/gpu_prof_int4.py sweep 2>&1 | tail -13
echo "=== flags 16 (N=32 single-buffered D, T=3)"; AO_B200_TS_FLAGS=16 timeout 150 python -u scripts/gpu_prof_int4.py sweep 2>&1 | tail -7
AO_B200_TS_FLAGS=16 timeout 200 python -u scripts/gpu_probe_int4.py --stage tc 2>&1 | grep -E "RESULT|FAIL|rror|identical|sqnr\(ours,fp32\)= *(-|nan|[0-3][0-9]\.)" | tail -4
echo "=== timeline"; timeout 100 python scripts/gpu_timeline.py 1,32 2>&1 | tail -16
