"""Setup-time quantization primitives (weights are quantized once, with plain torch ops on the
device, restating the reference's arithmetic; the per-call activation quantizers are CUDA kernels
in csrc/quant_act.cu).  Oracle-checked bit-exact in tests/test_primitives.py.

Reference: torchao/quantization/quant_primitives.py (_choose_qparams_affine_tinygemm :1268-1335,
_quantize_affine_tinygemm :488-599, _choose_qparams_affine :1487-1583, _quantize_affine_no_dtype_cast
:424-485, _choose_scale_float8 :2172-2212, _quantize_affine_float8 :2271-2287).
"""
from __future__ import annotations

from enum import Enum, auto
from typing import List, Optional, Tuple

import torch

__all__ = [
    "MappingType", "choose_qparams_affine_tinygemm", "quantize_affine_tinygemm", "choose_qparams_affine_int8",
    "quantize_affine_int8", "dequantize_affine_int8", "choose_scale_float8", "quantize_affine_float8",
    "dequantize_affine_float8",
]


class MappingType(Enum):
    SYMMETRIC = auto()
    SYMMETRIC_NO_CLIPPING_ERR = auto()
    ASYMMETRIC = auto()


def _group_view(t: torch.Tensor, block_size) -> Tuple[torch.Tensor, List[int]]:
    """view [..., d_i, ...] as [..., d_i/b_i, b_i, ...]; returns the view and the dims to reduce."""
    assert len(block_size) == t.dim(), f"block_size {block_size} vs tensor dim {t.dim()}"
    shape, red = [], []
    for d, b in zip(t.shape, block_size):
        b = d if b == -1 else b
        assert d % b == 0, f"dim {d} not divisible by block {b}"
        shape += [d // b, b]
        red.append(len(shape) - 1)
    return t.reshape(shape), red


# ---------------------------------------------------------------- int4 tinygemm (float zero point)
def choose_qparams_affine_tinygemm(w: torch.Tensor, group_size: int, quant_min=0, quant_max=15):
    """scale = (max-min)/15 clamp(eps), zero = min + 8*scale; every op in the input dtype (bf16),
    exactly as eager torch evaluates the reference.  Returns [N, K/g] each."""
    N, K = w.shape
    wg = w.reshape(N, K // group_size, group_size)
    mn = torch.amin(wg, dim=-1)
    mx = torch.amax(wg, dim=-1)
    eps = torch.finfo(w.dtype).smallest_normal
    scale = torch.clamp((mx - mn) / float(quant_max - quant_min), min=eps)
    mid = (quant_max + quant_min + 1) / 2
    zero = mn + scale * mid
    return scale.to(w.dtype), zero.to(w.dtype)


def quantize_affine_tinygemm(w, group_size, scale, zero, quant_min=0, quant_max=15) -> torch.Tensor:
    N, K = w.shape
    wg = w.reshape(N, K // group_size, group_size)
    s = scale.reshape(N, -1, 1)
    z = zero.reshape(N, -1, 1)
    mid = (quant_max + quant_min + 1) / 2
    min_val = z - s * mid
    q = torch.clamp(torch.round((wg - min_val) / s), quant_min, quant_max)
    return q.reshape(N, K).to(torch.int32)


# ---------------------------------------------------------------- int8 affine (integer zero point)
def choose_qparams_affine_int8(x: torch.Tensor, block_size, mapping_type=MappingType.SYMMETRIC,
                               quant_min=-128, quant_max=127, eps: Optional[float] = None):
    """keepdim=True semantics: scale/zero_point have x.ndim dims (reference Int8Tensor.from_hp)."""
    if eps is None:
        eps = torch.finfo(torch.float32).eps
    xv, red = _group_view(x, block_size)
    mn = torch.amin(xv, dim=red, keepdim=False)
    mx = torch.amax(xv, dim=red, keepdim=False)
    mn_neg = torch.min(mn, torch.zeros_like(mn))
    mx_pos = torch.max(mx, torch.zeros_like(mx))
    if mapping_type == MappingType.SYMMETRIC:
        amax = torch.max(-mn_neg, mx_pos)
        scale = amax / (float(quant_max - quant_min) / 2)
        scale = torch.clamp(scale, min=eps)
        zp = torch.full_like(scale, int((quant_max + quant_min + 1) / 2))
    elif mapping_type == MappingType.ASYMMETRIC:
        scale = (mx_pos - mn_neg) / float(quant_max - quant_min)
        scale = torch.clamp(scale, min=eps)
        zp = quant_min - torch.round(mn_neg / scale)
        zp = torch.clamp(zp, quant_min, quant_max)
    else:
        raise ValueError(f"unsupported mapping type {mapping_type}")
    out_shape = [d // (d if b == -1 else b) for d, b in zip(x.shape, block_size)]
    return scale.reshape(out_shape).to(torch.float32), zp.reshape(out_shape).to(torch.int8)


def quantize_affine_int8(x, block_size, scale, zero_point, quant_min=-128, quant_max=127) -> torch.Tensor:
    xv, red = _group_view(x, block_size)
    shape = list(xv.shape)
    for r in red:
        shape[r] = 1
    s = scale.reshape(shape)
    q = torch.round(xv * (1.0 / s))
    if zero_point is not None:
        q = q + zero_point.reshape(shape)
    return torch.clamp(q, quant_min, quant_max).reshape(x.shape).to(torch.int8)


def dequantize_affine_int8(q, block_size, scale, zero_point, output_dtype=torch.float32) -> torch.Tensor:
    qv, red = _group_view(q, block_size)
    shape = list(qv.shape)
    for r in red:
        shape[r] = 1
    v = qv.to(torch.int32)
    if zero_point is not None:
        v = v - zero_point.reshape(shape).to(torch.int32)
    return (v.to(scale.dtype) * scale.reshape(shape)).reshape(q.shape).to(output_dtype)


# ---------------------------------------------------------------- float8
def choose_scale_float8(x: torch.Tensor, block_size, float8_dtype=torch.float8_e4m3fn,
                        hp_value_lb=None, hp_value_ub=None) -> torch.Tensor:
    qmax = torch.finfo(float8_dtype).max
    xv, red = _group_view(x, block_size)
    amax = xv.abs().amax(dim=red, keepdim=False)
    if hp_value_lb is not None or hp_value_ub is not None:
        amax = torch.clamp(amax, min=hp_value_lb, max=hp_value_ub)
    scale = amax / qmax  # in the input dtype, then widened (reference has no eps)
    out_shape = [d // (d if b == -1 else b) for d, b in zip(x.shape, block_size)]
    return scale.reshape(out_shape).to(torch.float32)


def _expand_scale(scale, shape):
    if scale.numel() == 1 or all(a == b or a == 1 for a, b in zip(scale.shape, shape)):
        return scale
    out = scale
    for i, (t, s) in enumerate(zip(shape, scale.shape)):
        if t != s:
            out = out.repeat_interleave(t // s, dim=i)
    return out


def quantize_affine_float8(x, scale, float8_dtype=torch.float8_e4m3fn) -> torch.Tensor:
    qmax = torch.finfo(float8_dtype).max
    y = x.to(torch.float32) / _expand_scale(scale, x.shape)
    return y.clamp(min=-qmax, max=qmax).to(float8_dtype)


def dequantize_affine_float8(q, scale, output_dtype=torch.float32) -> torch.Tensor:
    return (q.to(torch.float32) * _expand_scale(scale, q.shape)).to(output_dtype)
