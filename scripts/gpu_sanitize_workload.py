"""Small invocations of every GEMM kernel + the HQQ solver, for compute-sanitizer (scripts/gpu_sanitize.sh)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ao_b200
ops = torch.ops.ao_b200
torch.manual_seed(0)
g = 32
for (M, N, K) in [(1, 256, 1024), (32, 512, 2048), (17, 136, 1024)]:
    qd = torch.randint(-2**31, 2**31 - 1, (N // 8, K // 128, 32, 4), device="cuda", dtype=torch.int32)
    sz = (torch.rand(K // g, N, 2, device="cuda") * 0.01).to(torch.bfloat16)
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    y = ops.int4_tilepacked_linear(x, qd, g, sz, None, N, 1)
    y2 = ops.int4_tilepacked_linear(x, qd, g, sz, None, N, 1)
    torch.cuda.synchronize()
    assert torch.equal(y, y2)
x = torch.randn(32, 1024, device="cuda").to(torch.bfloat16)
w = (torch.randn(256, 1024, device="cuda") * 0.05).to(torch.bfloat16)
xq, xs = ops.fp8_quantize_rowwise(x); wq, ws = ops.fp8_quantize_rowwise(w)
ops.fp8_rowwise_linear(xq, xs, wq, ws, None)
xq, xs = ops.int8_quantize_rowwise(x); wq, ws = ops.int8_quantize_rowwise(w)
ops.int8_dyn_linear(xq, xs, wq, ws, None)
ops.int4_hqq_quantize(w, 32)
torch.cuda.synchronize()
print("sanitize workload done")
