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
# round 2: the 128-token decode variant, the prefill-shaped kernel (needs >= 50 chunks of 128x256 per SM: a real-size
# GEMM), the nvfp4-weight pipelines and every quantizer
for (M, N, K) in [(100, 256, 1024), (128, 384, 2048), (512, 8192, 8192)]:
    qd = torch.randint(-2**31, 2**31 - 1, (N // 8, K // 128, 32, 4), device="cuda", dtype=torch.int32)
    sz = (torch.rand(K // g, N, 2, device="cuda") * 0.01).to(torch.bfloat16)
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    y = ops.int4_tilepacked_linear(x, qd, g, sz, None, N, 1)
    torch.cuda.synchronize()
for (M, N, K) in [(7, 256, 1024), (512, 8192, 8192)]:
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    pb = (w.float().abs().max() / (448.0 * 6.0)).reshape(1)
    wq, ws = ops.nvfp4_quantize(w, pb, True)
    xq, sx = ops.fp8_fakequant_rowwise(x)
    ops.nvfp4_weight_linear(xq, sx.reshape(-1), wq, ws, pb, None)
    torch.cuda.synchronize()
x = torch.randn(130, 4096, device="cuda").to(torch.bfloat16)
ops.mxfp8_quantize(x, True); ops.mxfp8_quantize(x, False); ops.nvfp4_quantize(x, None, True); ops.nvfp4_quantize(x, None, False)
ops.int8_quantize_rowwise(x); ops.fp8_quantize_rowwise(x)
xl = torch.randn(3, 32768, device="cuda").to(torch.bfloat16)
ops.int8_quantize_rowwise(xl); ops.fp8_quantize_rowwise(xl)
torch.cuda.synchronize()
x = torch.randn(32, 1024, device="cuda").to(torch.bfloat16)
w = (torch.randn(256, 1024, device="cuda") * 0.05).to(torch.bfloat16)
xq, xs = ops.fp8_quantize_rowwise(x); wq, ws = ops.fp8_quantize_rowwise(w)
ops.fp8_rowwise_linear(xq, xs, wq, ws, None)
xq, xs = ops.int8_quantize_rowwise(x); wq, ws = ops.int8_quantize_rowwise(w)
ops.int8_dyn_linear(xq, xs, wq, ws, None)
ops.int4_hqq_quantize(w, 32)
torch.cuda.synchronize()
print("sanitize workload done")
