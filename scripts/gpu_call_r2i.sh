#!/bin/bash
# quantizer rewrite (bit-exact tests + bandwidth), next-weight L2 prefetch A/B on the layer chain, eager CONTRIB publish
mkdir -p gpurun_out
echo "=== parity (quantizers, fusion, int4 decode)"; timeout 900 python -m pytest tests/test_lowp_gpu.py tests/test_fusion_gpu.py tests/test_int4_gpu.py tests/test_parity_holes_gpu.py -q -x 2>&1 | grep -v "^DEBUG\|^$\|Warning\|warnings.warn\|return Variable\|Consider using\|return float" | tail -6
echo "=== quant bw"; timeout 300 python -u scripts/gpu_quant_bw.py 2>&1 | tail -20
echo "=== layer chain: fused (no hint)"; timeout 300 python -u scripts/gpu_int4_layer.py one fused 2>&1 | tail -2
echo "=== layer chain: pf (every launch prefetches the next launch's weights into L2)"; timeout 300 python -u scripts/gpu_int4_layer.py one pf 2>&1 | tail -2
echo "--- pf, FLAGS=1 (no dequant arithmetic)"; AO_B200_TS_FLAGS=1 timeout 200 python -u scripts/gpu_int4_layer.py one pf 2>&1 | tail -2
echo "=== shapes"; timeout 300 python -u scripts/gpu_int4_layer.py shapes 2>&1 | tail -14
echo "=== bench"; timeout 900 python bench.py > gpurun_out/r02_bench_i.json 2> gpurun_out/r02_bench_i.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_i.json'))
print('bs32', d['ms_per_step'], d['roofline']['frac'], 'bs1', d['config']['bs1']['ms_per_step'], d['config']['bs1']['roofline_frac'], 'links', d['config'].get('next_weight_l2_prefetch_links'))
print('gpu_reference', d['gpu_reference']['bs32']['speedup'], d['gpu_reference']['bs1']['speedup'])
PY
tail -3 gpurun_out/r02_bench_i.err
