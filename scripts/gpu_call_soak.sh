#!/bin/bash
# soak: the whole GPU suite several times in a row (the round found two timing-dependent bugs; a flaky test must show here)
for rep in 1 2 3 4; do timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | grep -E "passed|failed|FAILED|Error" | head -4; done
echo "=== stress"; for g in o oac; do for M in 128 512 2048; do timeout 120 python -u scripts/gpu_stress_seq.py $M 6144 4096 7 $g 2>&1 | grep -E "ok$|Error" | tail -1; done; done
