"""Generate golden input/output fixtures by importing the REFERENCE (torchao 0.19) on CPU.

Run in the build container only (needs /root/reference):
    PYTHONPATH=/root/reference python tests/golden/make_golden.py
Writes small .npz files next to this script.  The fixtures pin oracle/ao_oracle.c (and through
it the CUDA path) to the reference's own arithmetic; /root/reference is never read at test time.

bf16 tensors are stored as uint16 bit patterns, fp8/e8m0/fp4 bytes as uint8.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("AO_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))


def bits(t):
    return t.detach().contiguous().view(torch.int16).numpy().view(np.uint16).copy()


def u8(t):
    return t.detach().contiguous().view(torch.uint8).numpy().copy()


def main():
    torch.manual_seed(1234)
    import torchao  # noqa: F401
    from torchao.quantization.quant_primitives import (
        MappingType,
        _choose_qparams_affine_tinygemm,
        _quantize_affine_tinygemm,
        _choose_scale_float8,
        _quantize_affine_float8,
    )
    from torchao.quantization.utils import pack_tinygemm_scales_and_zeros
    from torchao.quantization.granularity import PerRow
    from torchao.quantization.quantize_.workflows.int8.int8_tensor import Int8Tensor
    from torchao.prototype.mx_formats.mx_tensor import to_mx, ScaleCalculationMode, MXTensor
    from torchao.prototype.mx_formats.nvfp4_tensor import nvfp4_quantize, per_tensor_amax_to_scale, NVFP4Tensor
    from torchao.prototype.mx_formats.utils import to_blocked
    from torchao.prototype.mx_formats.kernels import f32_to_f4_unpacked, f4_unpacked_to_f32

    # ---------------- int4 tinygemm qparams / quantize (bf16 arithmetic) -------------------
    out = {}
    for g in (32, 128):
        N, K = 24, 256
        w = (torch.randn(N, K) * 0.02).to(torch.bfloat16)
        w[3, :g] = 0.0  # a constant group exercises the eps clamp
        w[5, 7] = 3.0   # outlier
        s, z = _choose_qparams_affine_tinygemm(
            w, MappingType.ASYMMETRIC, (1, g), torch.int32, 0, 15, scale_dtype=torch.bfloat16,
            zero_point_dtype=torch.bfloat16)
        q = _quantize_affine_tinygemm(w, [1, g], s, z, torch.int32, 0, 15)
        sz = pack_tinygemm_scales_and_zeros(s.reshape(N, -1), z.reshape(N, -1), torch.bfloat16)
        out[f"w_g{g}"] = bits(w)
        out[f"s_g{g}"] = bits(s.reshape(N, -1))
        out[f"z_g{g}"] = bits(z.reshape(N, -1))
        out[f"q_g{g}"] = q.to(torch.uint8).numpy()
        out[f"sz_g{g}"] = bits(sz)
    np.savez_compressed(os.path.join(HERE, "int4_tinygemm.npz"), **out)

    # ---------------- int4 CPU linear of the reference (BASELINE config[0] path) ------------
    # PrototypeInt4WeightOnlyConfig -> Int4OpaqueTensor -> aten._weight_int4pack_mm_for_cpu
    out = {}
    try:
        from torchao.quantization import quantize_
        from torchao.prototype.quantization.int4.int4_opaque_tensor import Int4OpaqueTensor  # noqa: F401
        from torchao.prototype.quantization import PrototypeInt4WeightOnlyConfig  # type: ignore
    except Exception:
        PrototypeInt4WeightOnlyConfig = None
        try:
            from torchao.prototype.quantization.int4.inference_workflow import PrototypeInt4WeightOnlyConfig  # type: ignore
        except Exception:
            pass
    if PrototypeInt4WeightOnlyConfig is not None:
        from torchao.quantization import quantize_
        N, K, M, g = 64, 256, 3, 32
        lin = torch.nn.Linear(K, N, bias=True).to(torch.bfloat16)
        w0 = lin.weight.detach().clone()
        b0 = lin.bias.detach().clone()
        x = torch.randn(M, K).to(torch.bfloat16)
        quantize_(lin, PrototypeInt4WeightOnlyConfig(group_size=g))
        y = lin(x)
        wt = lin.weight
        out["w"] = bits(w0)
        out["bias"] = bits(b0)
        out["x"] = bits(x)
        out["y"] = bits(y)
        out["sz"] = bits(wt.scale_and_zero)
        out["g"] = np.array(g)
        np.savez_compressed(os.path.join(HERE, "int4_cpu_linear.npz"), **out)
    else:
        print("WARNING: PrototypeInt4WeightOnlyConfig not importable; int4_cpu_linear.npz skipped")

    # ---------------- int8 per-row symmetric (weights and activations) + CPU linear ---------
    out = {}
    M, K, N = 5, 192, 40
    x = torch.randn(M, K).to(torch.bfloat16)
    x[2] = 0.0  # all-zero row exercises the eps clamp
    w = (torch.randn(N, K) * 0.05).to(torch.bfloat16)
    xt = Int8Tensor.from_hp(x, PerRow())
    wt = Int8Tensor.from_hp(w, PerRow())
    out["x"] = bits(x)
    out["w"] = bits(w)
    out["xq"] = xt.qdata.numpy()
    out["xs"] = xt.scale.reshape(-1).float().numpy()
    out["wq"] = wt.qdata.numpy()
    out["ws"] = wt.scale.reshape(-1).float().numpy()
    # the reference's own linear on CPU (dynamic act quant), bias included
    from torchao.quantization import quantize_, Int8DynamicActivationInt8WeightConfig
    lin = torch.nn.Linear(K, N, bias=True).to(torch.bfloat16)
    with torch.no_grad():
        lin.weight.copy_(w)
    b0 = lin.bias.detach().clone()
    quantize_(lin, Int8DynamicActivationInt8WeightConfig())
    y = lin(x)
    out["bias"] = bits(b0)
    out["y"] = bits(y)
    np.savez_compressed(os.path.join(HERE, "int8_rowwise.npz"), **out)

    # ---------------- fp8 e4m3 rowwise --------------------------------------------------------
    out = {}
    M, K = 6, 128
    x = (torch.randn(M, K) * 3).to(torch.bfloat16)
    x[1, 5] = 1000.0
    x[4] = x[4] * 1e-3
    s = _choose_scale_float8(x, [1, K], torch.float8_e4m3fn)
    q = _quantize_affine_float8(x, s, torch.float8_e4m3fn)
    out["x"] = bits(x)
    out["s"] = s.reshape(-1).numpy()
    out["q"] = u8(q)
    # decode table for all 256 e4m3 codes
    allc = torch.arange(256, dtype=torch.uint8).view(torch.float8_e4m3fn).float()
    out["e4m3_table"] = allc.numpy()
    # encode of a dense sweep (RNE incl. subnormals and the 448 edge)
    sweep = torch.cat([torch.linspace(-460, 460, 4001), torch.linspace(-0.05, 0.05, 2001)]).float()
    out["sweep"] = sweep.numpy()
    out["sweep_q"] = u8(sweep.clamp(-448, 448).to(torch.float8_e4m3fn))
    np.savez_compressed(os.path.join(HERE, "fp8_rowwise.npz"), **out)

    # ---------------- mxfp8 (RCEIL) -----------------------------------------------------------
    out = {}
    M, K = 130, 256
    x = (torch.randn(M, K) * torch.logspace(-3, 3, M).unsqueeze(1)).to(torch.bfloat16)
    x[7, :32] = 0.0
    sc, q = to_mx(x, torch.float8_e4m3fn, 32, ScaleCalculationMode.RCEIL, is_swizzled_scales=False)
    out["x"] = bits(x)
    out["scale"] = u8(sc)
    out["q"] = u8(q)
    sc_sw, _ = to_mx(x, torch.float8_e4m3fn, 32, ScaleCalculationMode.RCEIL, is_swizzled_scales=True)
    out["scale_swizzled"] = u8(sc_sw)
    out["blocked_of_plain"] = u8(to_blocked(sc.view(torch.uint8).reshape(M, K // 32)))
    mxt = MXTensor.to_mx(x, torch.float8_e4m3fn, 32, ScaleCalculationMode.RCEIL)
    out["dq"] = bits(mxt.dequantize(torch.bfloat16))
    # e8m0 rceil known answers incl. specials
    vals = torch.tensor([0.0, 1.0, 1.0000001, 0.75, 2.0, 3.0, 5.877e-39, 1e-45, 2.9e-39, 3.0e38,
                         float("inf"), float("nan"), 448.0, 1 / 448.0], dtype=torch.float32)
    from torchao.prototype.mx_formats.mx_tensor import _f32_to_e8m0_rceil
    out["e8m0_in"] = vals.numpy()
    out["e8m0_out"] = _f32_to_e8m0_rceil(vals).numpy()
    np.savez_compressed(os.path.join(HERE, "mxfp8.npz"), **out)

    # ---------------- nvfp4 ---------------------------------------------------------------------
    out = {}
    M, K = 9, 128
    x = (torch.randn(M, K) * 2).to(torch.bfloat16)
    x[0, :16] = 0.0
    x[3, 17] = 200.0
    sc1, q1 = nvfp4_quantize(x, 16, None)
    out["x"] = bits(x)
    out["scale_1lvl"] = u8(sc1)
    out["q_1lvl"] = q1.view(torch.uint8).numpy().copy()
    pts = per_tensor_amax_to_scale(x.abs().max().float())
    sc2, q2 = nvfp4_quantize(x, 16, pts)
    out["pts"] = pts.reshape(1).numpy()
    out["scale_2lvl"] = u8(sc2)
    out["q_2lvl"] = q2.view(torch.uint8).numpy().copy()
    t = NVFP4Tensor.to_nvfp4(x, per_tensor_scale=pts, is_swizzled_scales=False)
    out["dq_2lvl"] = bits(t.dequantize(torch.bfloat16))
    out["dq_2lvl_f32"] = t.dequantize(torch.float32).numpy()
    t1 = NVFP4Tensor.to_nvfp4(x, is_swizzled_scales=False)
    out["dq_1lvl"] = bits(t1.dequantize(torch.bfloat16))
    # e2m1 encode sweep + decode table
    sweep = torch.linspace(-7, 7, 2801).float()
    out["f4_sweep"] = sweep.numpy()
    out["f4_sweep_q"] = f32_to_f4_unpacked(sweep).numpy()
    out["f4_table"] = f4_unpacked_to_f32(torch.arange(16, dtype=torch.uint8)).numpy()
    # blocked layout of an odd-shaped scale matrix
    sm = torch.arange(200 * 6, dtype=torch.int32).remainder(251).to(torch.uint8).reshape(200, 6)
    out["blk_in"] = sm.numpy()
    out["blk_out"] = to_blocked(sm).numpy()
    np.savez_compressed(os.path.join(HERE, "nvfp4.npz"), **out)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
