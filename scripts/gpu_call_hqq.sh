#!/bin/bash
timeout 300 python -m pytest tests/test_int4_gpu.py -q -m gpu -k "hqq" --timeout 90 2>&1 | tail -15
timeout 100 python - <<'PY'
import torch, time, ao_b200
w=(torch.randn(14336,4096,device="cuda")*0.02).to(torch.bfloat16)
for _ in range(2): torch.ops.ao_b200.int4_hqq_quantize(w,32)
torch.cuda.synchronize(); t=time.time()
for _ in range(5): torch.ops.ao_b200.int4_hqq_quantize(w,32)
torch.cuda.synchronize(); print("hqq 14336x4096 g=32: %.2f ms per weight"%((time.time()-t)/5*1e3))
PY
