#!/bin/bash
# early stage release + hoisted epilogue bookkeeping (layer chain, shapes), per-shape lowp GEMMs vs the library kernels
mkdir -p gpurun_out
echo "=== parity"; timeout 900 python -m pytest tests/test_int4_gpu.py tests/test_fusion_gpu.py -q -x 2>&1 | grep -v "^DEBUG\|^$\|Warning\|warnings.warn\|return Variable\|Consider using\|return float" | tail -3
echo "=== layer chain"; timeout 300 python -u scripts/gpu_int4_layer.py one fused 2>&1 | tail -2
echo "=== shapes"; timeout 300 python -u scripts/gpu_int4_layer.py shapes 2>&1 | tail -12
echo "=== lowp vs library"; timeout 1200 python -u scripts/gpu_lowp_vs_library.py 2>&1 | tail -70
