#!/bin/bash
# quantizer rework (parity + GB/s), per-line L2 prefetch variants on the layer chain, ncu captures (decode + prefill kernels)
mkdir -p gpurun_out
echo "=== parity (quantizers)"; timeout 900 python -m pytest tests/test_lowp_gpu.py tests/test_fusion_gpu.py tests/test_parity_holes_gpu.py -q -x 2>&1 | grep -v "^DEBUG\|^$\|Warning\|warnings.warn\|return Variable\|Consider using\|return float" | tail -4
echo "=== quant bw"; timeout 300 python -u scripts/gpu_quant_bw.py 2>&1 | head -13
echo "=== layer chain: fused (no hint)"; timeout 300 python -u scripts/gpu_int4_layer.py one fused 2>&1 | tail -2
for mode in 4 5 6; do for cap in 0 16; do
  echo "--- pf mode $mode cap ${cap} MB"; AO_B200_PF_MODE=$mode AO_B200_PF_CAP_MB=$cap timeout 200 python -u scripts/gpu_int4_layer.py one pf 2>&1 | tail -2
done; done
echo "=== ncu decode"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:ts_gemm -c 8 -o gpurun_out/r02_int4_decode -f python scripts/gpu_ncu_shapes_r2.py decode > gpurun_out/ncu_decode.log 2>&1; tail -2 gpurun_out/ncu_decode.log
echo "=== ncu prefill"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:ts_prefill -c 4 -o gpurun_out/r02_int4_prefill -f python scripts/gpu_ncu_shapes_r2.py prefill > gpurun_out/ncu_prefill.log 2>&1; tail -2 gpurun_out/ncu_prefill.log
ls -la gpurun_out/*.ncu-rep
