#!/bin/bash
# round-end style sequence on the current tree: full GPU suite, bench line, launch list of a bench step, smoke()
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^DEBUG\|^$\|Warning\|warnings.warn\|return Variable\|Consider using\|return float" | tail -4
echo "=== bench"; timeout 1200 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -3 gpurun_out/r02_bench_final.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_final.json'))
print('bs32', d['ms_per_step'], d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'bs1', d['config']['bs1']['ms_per_step'], d['config']['bs1']['roofline_frac'])
print('gpu_reference', d['gpu_reference']['bs32']['speedup'], d['gpu_reference']['bs1']['speedup'])
for k,v in d['configs'].items():
    for kk,vv in v.items():
        if isinstance(vv,dict) and 'ms_per_step' in vv: print(k,kk,round(vv['ms_per_step'],3),round(vv['roofline_frac'],3),vv.get('library',{}).get('speedup'))
for k,v in d.get('prefill',{}).items():
    if isinstance(v,dict): print(k, {f:(round(r['tflops']),round(r['frac'],3), round(r.get('library_us',0),1)) for f,r in v.items() if isinstance(r,dict)}, 'bf16 us', round(v['bf16_cublas_us'],1))
PY
echo "=== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ts_gemm -c 300 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 1 --quick > gpurun_out/bench_under_ncu.log 2>&1; wc -l gpurun_out/r02_bench_launches.csv
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
