"""int8 per-row weight with dynamic per-token int8 activations on tcgen05 kind::i8.

Attribute names / order match torchao's Int8Tensor
(torchao/quantization/quantize_/workflows/int8/int8_tensor.py:74-85).  The linear replaces
``Int8Tensor.from_hp(x) -> _int_scaled_matmul -> aten._int_mm -> pointwise epilogue``
(:266-359, int8/kernels.py:114-144) with two launches: a fused per-token quantizer and an int8 GEMM
whose epilogue reproduces the reference's rounding order bf16(bf16(acc*s_x)*s_w + bias).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch

from ao_b200.quantization.granularity import Granularity, PerRow, PerTensor
from ao_b200.quantization.quant_primitives import (
    MappingType, choose_qparams_affine_int8, dequantize_affine_int8, quantize_affine_int8)
from ao_b200.quantization.quantize_.common.quantize_tensor_kwargs import QuantizeTensorKwargs
from ao_b200.quantization.utils import get_block_size
from ao_b200.utils import TorchAOBaseTensor, fill_defaults, rows_for_kernel

__all__ = ["Int8Tensor", "QuantizeTensorToInt8Kwargs"]
aten = torch.ops.aten


@dataclass
class QuantizeTensorToInt8Kwargs(QuantizeTensorKwargs):
    granularity: Granularity
    mapping_type: MappingType = MappingType.SYMMETRIC
    reduce_range: bool = False


class Int8Tensor(TorchAOBaseTensor):
    tensor_data_names = ["qdata", "scale"]
    optional_tensor_data_names = ["zero_point", "act_quant_scale", "act_quant_zero_point", "act_pre_scale"]
    tensor_attribute_names = ["block_size", "dtype"]
    optional_tensor_attribute_names = ["act_quant_kwargs", "reduce_range"]

    def __new__(cls, qdata, scale, block_size: List[int], dtype: torch.dtype, zero_point=None, act_quant_scale=None,
                act_quant_zero_point=None, act_pre_scale=None, act_quant_kwargs=None, reduce_range=False):
        return torch.Tensor._make_wrapper_subclass(cls, qdata.shape, device=qdata.device, dtype=dtype, requires_grad=False)

    def __init__(self, qdata, scale, block_size, dtype, zero_point=None, act_quant_scale=None,
                 act_quant_zero_point=None, act_pre_scale=None, act_quant_kwargs=None, reduce_range=False):
        super().__init__()
        self.qdata = qdata
        self.scale = scale
        self.block_size = block_size
        self.zero_point = zero_point
        self.act_quant_scale = act_quant_scale
        self.act_quant_zero_point = act_quant_zero_point
        self.act_pre_scale = act_pre_scale
        self.act_quant_kwargs = act_quant_kwargs
        self.reduce_range = reduce_range

    def _quantization_type(self):
        return f"{self.act_quant_kwargs=}, {self.block_size=}, {self.scale.shape=}"

    @staticmethod
    def _normalize_granularity(granularity):
        if granularity is None:
            return PerRow(), PerRow()
        if isinstance(granularity, (list, tuple)):
            assert len(granularity) == 2, "granularity list must be [activation, weight]"
            return granularity[0], granularity[1]
        return granularity, granularity

    @classmethod
    def from_hp(cls, hp_tensor: torch.Tensor, granularity: Granularity, mapping_type=MappingType.SYMMETRIC,
                scale=None, zero_point=None, act_quant_kwargs=None, act_quant_scale=None,
                act_quant_zero_point=None, act_pre_scale=None, reduce_range: Optional[bool] = False):
        block_size = list(get_block_size(hp_tensor.shape, granularity))
        qmin, qmax = (-64, 63) if reduce_range else (-128, 127)
        fast = (scale is None and hp_tensor.is_cuda and hp_tensor.dtype == torch.bfloat16
                and mapping_type == MappingType.SYMMETRIC and isinstance(granularity, PerRow)
                and granularity.dim in (-1, hp_tensor.dim() - 1) and not reduce_range and hp_tensor.shape[-1] % 8 == 0)
        if fast:
            x2 = rows_for_kernel(hp_tensor.reshape(-1, hp_tensor.shape[-1]))
            q, s = torch.ops.ao_b200.int8_quantize_rowwise(x2)
            int_data = q.reshape(hp_tensor.shape)
            scale = s.reshape(*hp_tensor.shape[:-1], 1)
            zero_point = torch.zeros_like(scale, dtype=torch.int8)
        else:
            if scale is None:
                scale, zero_point = choose_qparams_affine_int8(hp_tensor, block_size, mapping_type, qmin, qmax)
            else:
                assert scale.ndim == hp_tensor.ndim
                if zero_point is None:
                    zero_point = torch.zeros_like(scale, dtype=torch.int8)
            int_data = quantize_affine_int8(hp_tensor, block_size, scale, zero_point, qmin, qmax)
        return cls(int_data, scale, block_size, hp_tensor.dtype, zero_point=zero_point, act_quant_scale=act_quant_scale,
                   act_quant_zero_point=act_quant_zero_point, act_pre_scale=act_pre_scale,
                   act_quant_kwargs=act_quant_kwargs, reduce_range=reduce_range)

    def dequantize(self, output_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        return dequantize_affine_int8(self.qdata, self.block_size, self.scale, self.zero_point,
                                      output_dtype if output_dtype is not None else self.dtype)


implements = Int8Tensor.implements
implements_torch_function = Int8Tensor.implements_torch_function


@implements(aten.linear.default)
@implements_torch_function(torch.nn.functional.linear)
def _(func, types, args, kwargs):
    x, w, bias = args[0], args[1], args[2] if len(args) > 2 else None
    assert isinstance(w, Int8Tensor), f"Expected weight to be Int8Tensor, got {type(w)}"
    out_dtype = x.dtype
    if w.act_pre_scale is not None:
        x = x * w.act_pre_scale
    if w.act_quant_kwargs is None:
        raise NotImplementedError(
            "int8 weight-only linear is outside this engine's scope (SURVEY §8); use "
            "Int8DynamicActivationInt8WeightConfig")
    N, K = w.qdata.shape[-2], w.qdata.shape[-1]
    x2 = x.reshape(-1, K)
    out_shape = (*x.shape[:-1], N)
    if x2.shape[0] == 0:
        return x.new_empty(out_shape)
    kw = w.act_quant_kwargs
    xt = Int8Tensor.from_hp(x2, kw.granularity, mapping_type=kw.mapping_type, scale=w.act_quant_scale,
                            zero_point=w.act_quant_zero_point, reduce_range=kw.reduce_range)
    M = x2.shape[0]
    xs = xt.scale.reshape(-1).float()
    if xs.numel() == 1:
        xs = xs.expand(M)
    ws = w.scale.reshape(-1).float()
    if ws.numel() == 1:
        ws = ws.expand(N)
    wq = w.qdata.contiguous()
    if kw.mapping_type == MappingType.SYMMETRIC:
        y = torch.ops.ao_b200.int8_dyn_linear(xt.qdata.contiguous(), xs.contiguous(), wq, ws.contiguous(), bias)
        return y.reshape(out_shape).to(out_dtype)
    # asymmetric activations (int8_tensor.py:322-330): Y = (Xq Wq^T) s_x s_w - zp_x s_x rowsum(Wq) s_w
    acc = torch.ops.ao_b200.int8_mm_i32(xt.qdata.contiguous(), wq)
    y = (acc * xs.reshape(-1, 1)).to(out_dtype)
    zp = xt.zero_point.reshape(-1, 1).float()
    corr = (zp * xs.reshape(-1, 1)) * w.qdata.sum(dim=-1).float()
    y = (y - corr.to(out_dtype)) * ws
    if bias is not None:
        y = y + bias
    return y.reshape(out_shape).to(out_dtype)


@implements(aten.slice.Tensor)
def _(func, types, args, kwargs):
    self, dim, start, end, step = fill_defaults(args, 5, [0, None, None, 1])
    assert step == 1 and dim in (0, 1, 2)
    qd = aten.slice.Tensor(self.qdata, dim, start, end, step)
    sc, zp = self.scale, self.zero_point
    if sc.numel() > 1 and sc.dim() == self.qdata.dim() and sc.shape[dim] == self.qdata.shape[dim]:
        sc = aten.slice.Tensor(sc, dim, start, end, step)
        if zp is not None:
            zp = aten.slice.Tensor(zp, dim, start, end, step)
    bs = list(self.block_size)
    if bs[dim] > qd.shape[dim]:
        bs[dim] = qd.shape[dim]
    return Int8Tensor(qd, sc, bs, self.dtype, zero_point=zp, act_quant_scale=self.act_quant_scale,
                      act_quant_zero_point=self.act_quant_zero_point, act_pre_scale=self.act_pre_scale,
                      act_quant_kwargs=self.act_quant_kwargs, reduce_range=self.reduce_range)


@implements(aten.select.int)
def _(func, types, args, kwargs):
    self, dim, index = args
    assert dim == 0
    sc = self.scale[index] if self.scale.dim() == self.qdata.dim() and self.scale.shape[0] == self.qdata.shape[0] else self.scale
    zp = self.zero_point
    if zp is not None and zp.dim() == self.qdata.dim() and zp.shape[0] == self.qdata.shape[0]:
        zp = zp[index]
    return Int8Tensor(self.qdata[index], sc, list(self.block_size[1:]), self.dtype, zero_point=zp,
                      act_quant_kwargs=self.act_quant_kwargs, reduce_range=self.reduce_range)


Int8Tensor.__module__ = "ao_b200.quantization"
torch.serialization.add_safe_globals([Int8Tensor, QuantizeTensorToInt8Kwargs, MappingType])
