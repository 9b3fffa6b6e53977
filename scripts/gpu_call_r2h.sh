#!/bin/bash
# prefill kernel bring-up: parity tests, A/B against the 128-token-block path, the launch failure of call f
mkdir -p gpurun_out
echo "=== prefill parity"; timeout 900 python -m pytest tests/test_int4_gpu.py tests/test_lowp_gpu.py -q -x -k "prefill or nvfp4_weight" 2>&1 | grep -v "^DEBUG\|^$\|Warning\|warnings.warn\|return Variable\|Consider using\|return float" | tail -6
echo "=== prefill kernel (256-token tiles)"; timeout 600 python -u scripts/gpu_prefill.py int4 2>&1 | tail -12
echo "=== old path, per process"
for s in 6144x4096 4096x4096 28672x4096 4096x14336; do AO_B200_NO_PREFILL=1 timeout 300 python -u scripts/gpu_prefill.py int4 512,4096 $s 2>&1 | tail -3; done
echo "=== old path, one process (the sequence that failed in call f)"; AO_B200_NO_PREFILL=1 timeout 300 python -u scripts/gpu_prefill.py int4 512 2>&1 | tail -6
echo "=== same, no PDL"; AO_B200_NO_PREFILL=1 AO_B200_NO_PDL=1 timeout 300 python -u scripts/gpu_prefill.py int4 512 2>&1 | tail -6
echo "=== fp8 / int8 at prefill shapes"; timeout 600 python -u scripts/gpu_prefill.py fp8 2>&1 | tail -9; timeout 600 python -u scripts/gpu_prefill.py int8 2>&1 | tail -9
