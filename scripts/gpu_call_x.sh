#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_int4_gpu.py tests/test_lowp_gpu.py -q -m gpu -x --timeout 60 2>&1 | tail -2
timeout 100 python -u scripts/gpu_probe_int4.py --stages tc 2>&1 | grep -E "RESULT|FAIL|rror|identical" | tail -3
echo "=== sweep"; timeout 150 python -u scripts/gpu_prof_int4.py sweep 2>&1 | tail -12
echo "=== bench"; timeout 400 python bench.py --steps 20 --warmup 3 2> gpurun_out/bench_x.err | tee gpurun_out/bench_x.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bs32', d['ms_per_step'], d['value'], 'bs1', d['bs1'])"
