#!/bin/bash
# where does the 256-token prefill kernel beat the 128-token-block path?  (both correct now)  + lowp table with the new grid rule
mkdir -p gpurun_out
echo "=== prefill kernel"; timeout 900 python -u scripts/gpu_prefill.py int4 256,512,1024,2048,4096 2>&1 | cut -c1-75 | tail -22
echo "=== 128-token-block path"; AO_B200_NO_PREFILL=1 timeout 900 python -u scripts/gpu_prefill.py int4 256,512,1024,2048,4096 2>&1 | cut -c1-75 | tail -22
echo "=== X loads at bs=32 (FLAGS=4: none after the first ring-full)"; timeout 200 python -u scripts/gpu_int4_layer.py one fused 32 2>&1 | tail -1; AO_B200_TS_FLAGS=4 timeout 200 python -u scripts/gpu_int4_layer.py one fused 32 2>&1 | tail -1
echo "=== lowp vs library (min 8 chunks per CTA)"; timeout 1200 python -u scripts/gpu_lowp_vs_library.py 2>&1 | tail -62
