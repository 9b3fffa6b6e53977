#!/bin/bash
# re-entry baseline: full GPU test suite, prefill-shaped runs, quantizer bandwidth, small-shape timelines, bench line
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^DEBUG\|^$\|Warning\|warnings.warn\|return Variable\|Consider using\|return float" | tail -5
echo "=== prefill"; timeout 600 python -u scripts/gpu_prefill.py all 2>&1 | tail -30
echo "=== quant bw"; timeout 300 python -u scripts/gpu_quant_bw.py 2>&1 | tail -20
echo "=== timeline"; for s in 4096x4096 6144x4096; do timeout 120 python -u scripts/gpu_timeline.py 1,32 $s 2>&1 | tail -28; done
echo "=== bench"; timeout 900 python bench.py > gpurun_out/r02_bench_f.json 2> gpurun_out/r02_bench_f.err; tail -c 600 gpurun_out/r02_bench_f.json; tail -3 gpurun_out/r02_bench_f.err
