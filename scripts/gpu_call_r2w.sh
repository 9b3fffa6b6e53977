#!/bin/bash
# 2-GPU run exactly as the driver launches it (torchrun, NCCL), quick variant of both arms
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; echo "rc=$?"; tail -3 gpurun_out/r02_bench_2gpu.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_2gpu.json'))
print('n_gpus', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'bcast', d['config']['weight_broadcast_bytes'])
for k,v in d.get('configs',{}).items():
    for kk,vv in v.items():
        if isinstance(vv,dict) and 'ms_per_step' in vv: print(k,kk,round(vv['ms_per_step'],3),round(vv['value'],1))
PY
echo "=== reference arm under torchrun"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | cut -c1-300
