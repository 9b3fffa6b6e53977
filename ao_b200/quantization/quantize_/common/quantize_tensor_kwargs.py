"""Activation-quantization recipe attached to a weight tensor
(reference: torchao/quantization/quantize_/common/quantize_tensor_kwargs.py:36-71)."""
import abc


class QuantizeTensorKwargs(abc.ABC):
    """Base class for the kwargs dataclasses each tensor type defines for dynamic activation quant."""


def _choose_quant_func_and_quantize_tensor(tensor, quant_kwargs, **extra):
    from ao_b200.quantization.quantize_.workflows.float8.float8_tensor import Float8Tensor, QuantizeTensorToFloat8Kwargs
    from ao_b200.quantization.quantize_.workflows.int8.int8_tensor import Int8Tensor, QuantizeTensorToInt8Kwargs

    if isinstance(quant_kwargs, QuantizeTensorToFloat8Kwargs):
        return Float8Tensor.from_hp(tensor, quant_kwargs.float8_dtype, quant_kwargs.granularity, quant_kwargs.mm_config,
                                    quant_kwargs.hp_value_lb, quant_kwargs.hp_value_ub, quant_kwargs.kernel_preference)
    if isinstance(quant_kwargs, QuantizeTensorToInt8Kwargs):
        return Int8Tensor.from_hp(tensor, quant_kwargs.granularity, mapping_type=quant_kwargs.mapping_type,
                                  scale=extra.get("scale"), zero_point=extra.get("zero_point"))
    raise NotImplementedError(f"Quant kwargs not supported: {quant_kwargs}")
