#!/bin/bash
# final bench line of the round on 1 GPU (full) -- the 2-GPU run is a separate gpurun --gpus 2 call
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/r02_bench_u.json 2> gpurun_out/r02_bench_u.err; tail -2 gpurun_out/r02_bench_u.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_u.json'))
print('bs32', d['ms_per_step'], d['value'], d['roofline']['frac'], 'e2e', d['e2e']['value'], 'bs1', d['config']['bs1']['ms_per_step'], d['config']['bs1']['roofline_frac'])
print('gpu_reference', d['gpu_reference']['bs32']['speedup'], d['gpu_reference']['bs1']['speedup'], 'cpu', d['cpu_baseline']['value'])
for k,v in d['configs'].items():
    for kk,vv in v.items():
        if isinstance(vv,dict) and 'ms_per_step' in vv: print(k,kk,round(vv['ms_per_step'],3),round(vv['roofline_frac'],3),vv.get('library',{}).get('speedup'))
for k,v in d.get('prefill',{}).items():
    if isinstance(v,dict): print(k, {f:(round(r['tflops']),round(r['frac'],3), round(r.get('library_us',0),1)) for f,r in v.items() if isinstance(r,dict)}, 'bf16 us', round(v['bf16_cublas_us'],1))
PY
echo "=== reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-600
