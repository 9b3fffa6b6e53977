#!/bin/bash
# after the weight-stage parity fix: stress of the 128-token variant, full GPU suite, prefill TFLOP/s, quantizer GB/s,
# next-weight L2 prefetch timing modes on the layer chain
mkdir -p gpurun_out
run() { echo "--- $*"; timeout 120 python -u scripts/gpu_stress_seq.py "$@" 2>&1 | grep -E "ok$|Error|error" | tr '\n' ';' | cut -c1-300; echo; }
echo "=== stress (failed 9 of 12 times before the fix)"; export AO_B200_NO_PREFILL=1; for g in oac oa oc o o o; do run 128 6144 4096 7 $g; done; unset AO_B200_NO_PREFILL
for g in oac o o; do run 512 6144 4096 7 $g; done
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^DEBUG\|^$\|Warning\|warnings.warn\|return Variable\|Consider using\|return float" | tail -6
echo "=== prefill kernel"; timeout 600 python -u scripts/gpu_prefill.py int4 2>&1 | tail -10
echo "=== quant bw"; timeout 300 python -u scripts/gpu_quant_bw.py 2>&1 | head -13
echo "=== layer chain: fused (no hint)"; timeout 300 python -u scripts/gpu_int4_layer.py one fused 2>&1 | tail -2
for mode in 1 2 3; do for cap in 0 8 24; do
  echo "--- pf mode $mode cap ${cap} MB"; AO_B200_PF_MODE=$mode AO_B200_PF_CAP_MB=$cap timeout 200 python -u scripts/gpu_int4_layer.py one pf 2>&1 | tail -2
done; done
