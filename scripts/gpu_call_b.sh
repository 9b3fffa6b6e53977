#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/gpu_probe_int4.py --stages diag,tc --timeout 90 > gpurun_out/v3_probe.log 2>&1; grep -E "RESULT|mismatch|SUMMARY|sqnr\(ours,fp32\)= *[0-4][0-9]\." gpurun_out/v3_probe.log | tail -8
timeout 120 python scripts/gpu_timeline.py > gpurun_out/v3_timeline.log 2>&1; cat gpurun_out/v3_timeline.log | tail -20
timeout 200 python scripts/gpu_prof_int4.py sweep > gpurun_out/v3_sweep.log 2>&1; cat gpurun_out/v3_sweep.log
timeout 300 python scripts/gpu_probe_lowp.py --stages int8,fp8,mxfp8,bench --timeout 100 > gpurun_out/lowp2.log 2>&1; grep -E "RESULT|bench|FAIL|failed" gpurun_out/lowp2.log | tail -40
