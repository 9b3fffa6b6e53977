#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/gpu_probe_int4.py --stages diag,tc --timeout 90 > gpurun_out/v4_probe.log 2>&1; grep -E "RESULT|mismatch|SUMMARY|TIMEOUT|rror" gpurun_out/v4_probe.log | tail -8
timeout 120 python scripts/gpu_timeline.py > gpurun_out/v4_timeline.log 2>&1; cat gpurun_out/v4_timeline.log | tail -24
timeout 200 python scripts/gpu_prof_int4.py sweep > gpurun_out/v4_sweep.log 2>&1; cat gpurun_out/v4_sweep.log
