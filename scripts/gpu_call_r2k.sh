#!/bin/bash
# (1) which launch group faults for the N_MMA = 128 variant; (2) one-hot failure of the prefill kernel at M = 300
run() { echo "--- $*"; timeout 120 python -u scripts/gpu_stress_seq.py "$@" 2>&1 | grep -E "ok$|Error|error" | tr '\n' ';' | cut -c1-600; echo; }
export AO_B200_NO_PREFILL=1
for g in oac oac oa oa oc oc ac o; do run 128 6144 4096 7 $g; done
for g in oac oa oc; do run 64 6144 4096 7 $g; done
echo "=== no PDL"; export AO_B200_NO_PDL=1; for g in oac oac oa oc; do run 128 6144 4096 7 $g; done; unset AO_B200_NO_PDL
unset AO_B200_NO_PREFILL
echo "=== one-hot, prefill kernel"; for s in "300 6144 4096" "512 6144 4096" "300 4096 4096"; do timeout 120 python -u scripts/gpu_onehot_debug.py $s 2>&1 | tail -12; done
echo "=== one-hot, old path"; AO_B200_NO_PREFILL=1 timeout 120 python -u scripts/gpu_onehot_debug.py 300 6144 4096 2>&1 | tail -12
