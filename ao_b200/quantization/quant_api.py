"""User API: ``quantize_`` + the stable configs + their module handlers.

Same names, dataclass fields and defaults as torchao/quantization/quant_api.py
(quantize_ :249-321, Int4WeightOnlyConfig :502-535, Int8DynamicActivationInt8WeightConfig :808-862,
Float8DynamicActivationFloat8WeightConfig :1112-1172, FqnToConfig :1514-1600; handlers :597-627,
:885-915, :1299-1336).  Differences, all deliberate:
  * handlers do NOT touch global inductor flags (the reference calls
    recommended_inductor_config_setter(); this engine has no compiler in the hot path) --
    ``set_inductor_config`` is kept as an accepted, ignored field;
  * int4 packing formats other than TILE_PACKED_TO_4D raise (they need ``mslk`` in the reference too).
"""
from __future__ import annotations

import logging
import re
import types
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Any, Callable, List, Optional, Tuple, Union

import torch
import torch.nn as nn

from ao_b200._native import require_sm100
from ao_b200.core.config import AOBaseConfig
from ao_b200.float8.inference import FP8Granularity, Float8MMConfig, _normalize_granularity
from ao_b200.quantization.granularity import Granularity, PerRow, PerTensor
from ao_b200.quantization.quant_primitives import MappingType
from ao_b200.quantization.quantize_.common.kernel_preference import KernelPreference
from ao_b200.quantization.quantize_.workflows import (
    Float8PackingFormat, Float8Tensor, Int4ChooseQParamsAlgorithm, Int4PackingFormat, Int4TilePackedTo4dTensor,
    Int8Tensor, QuantizeTensorToFloat8Kwargs, QuantizeTensorToInt8Kwargs)
from ao_b200.quantization.transform_module import _QUANTIZE_CONFIG_HANDLER, register_quantize_module_handler

logger = logging.getLogger(__name__)

__all__ = [
    "quantize_", "Int4WeightOnlyConfig", "Int8DynamicActivationInt8WeightConfig",
    "Float8DynamicActivationFloat8WeightConfig", "FqnToConfig", "ModuleFqnToConfig", "fqn_matches_fqn_config",
    "_is_linear", "_replace_with_custom_fn_if_matches_filter",
]


# ---------------------------------------------------------------------------------------------
# module walk
# ---------------------------------------------------------------------------------------------
def _replace_with_custom_fn_if_matches_filter(model, replacement_fn, filter_fn, cur_fqn="", device=None,
                                              extra_args: Optional[Tuple[Any, ...]] = ()):
    """Depth-first: replace every child for which filter_fn(child, fqn) holds (reference :120-163)."""
    if filter_fn(model, cur_fqn[:-1]):
        if device is not None:
            model.to(device=device)
        return replacement_fn(model, *extra_args)
    for name, child in list(model.named_children()):
        new_child = _replace_with_custom_fn_if_matches_filter(child, replacement_fn, filter_fn, f"{cur_fqn}{name}.",
                                                              device, extra_args)
        if new_child is not child and new_child is not None:
            setattr(model, name, new_child)
    if device is not None:
        model.to(device=device)
    return model


def _is_linear(mod, *args):
    return (isinstance(mod, torch.nn.Linear) and hasattr(mod, "weight")
            and not isinstance(mod, nn.modules.linear.NonDynamicallyQuantizableLinear))


def _linear_extra_repr(self):
    from ao_b200.utils import TorchAOBaseTensor

    w = self.weight
    desc = f"{type(w).__name__}({w._quantization_type()})" if isinstance(w, TorchAOBaseTensor) and hasattr(w, "_quantization_type") else "not quantized"
    return f"in_features={w.shape[1]}, out_features={w.shape[0]}, weight={desc}"


def _set_quantized_param(module, parameter_name, new_tensor):
    setattr(module, parameter_name, torch.nn.Parameter(new_tensor, requires_grad=False))
    module.extra_repr = types.MethodType(_linear_extra_repr, module)
    return module


def quantize_(model: torch.nn.Module, config: AOBaseConfig,
              filter_fn: Optional[Callable[[torch.nn.Module, str], bool]] = _is_linear,
              device: Optional[torch.types.Device] = None):
    """Convert the weight of linear modules in ``model`` according to ``config``, in place; returns None."""
    if isinstance(config, FqnToConfig):
        if filter_fn is not None:   # the default `_is_linear` included, exactly like the reference (quant_api.py:286-290)
            raise ValueError("Custom filter_fn and FqnToConfig were both specified. Only filter_fn=None is supported "
                             "when FqnToConfig is specified.")
        named_modules = dict(model.named_modules())
        for module_fqn, module in named_modules.items():
            if (fqn_matches_fqn_config(module_fqn, config) or _module_param_matches_fqn_config(module, module_fqn, config)
                    or ("_default" in config.fqn_to_config and _is_linear(module))):
                replacement = _fqn_to_config_handler(module, module_fqn, config)
                if device is not None:
                    replacement = replacement.to(device=device)
                if replacement is not module and module_fqn != "":
                    child = module_fqn.split(".")[-1]
                    parent = named_modules[module_fqn.removesuffix(child).removesuffix(".")]
                    setattr(parent, child, replacement)
        return
    if not isinstance(config, AOBaseConfig):
        raise AssertionError("quantize_ expects an AOBaseConfig workflow configuration object")
    filter_fn = _is_linear if filter_fn is None else filter_fn
    if type(config) not in _QUANTIZE_CONFIG_HANDLER:
        raise KeyError(f"no quantize handler registered for {type(config).__name__}")
    handler = _QUANTIZE_CONFIG_HANDLER[type(config)]
    _replace_with_custom_fn_if_matches_filter(model, handler, filter_fn, device=device, extra_args=(config,))


# ---------------------------------------------------------------------------------------------
# int4 weight only
# ---------------------------------------------------------------------------------------------
@dataclass
class Int4WeightOnlyConfig(AOBaseConfig):
    """int4 group-wise weight-only quantization.  ``group_size`` in {256,128,64,32}.
    BASELINE uses ``int4_packing_format="tile_packed_to_4d"`` and ``group_size=32``."""

    group_size: int = 128
    set_inductor_config: bool = True
    int4_packing_format: Int4PackingFormat = Int4PackingFormat.PLAIN
    int4_choose_qparams_algorithm: Int4ChooseQParamsAlgorithm = Int4ChooseQParamsAlgorithm.TINYGEMM
    int4_tile_packed_ntile: int = 8
    version: int = 2

    def __post_init__(self):
        assert self.int4_tile_packed_ntile in [8, 16], "int4_tile_packed_ntile must be either 8 or 16"
        # like the reference, the two format fields keep what the caller passed (a plain string stays a string in
        # the config JSON); both are str-Enums, so comparisons with the enum members work either way
        Int4PackingFormat(self.int4_packing_format)                      # validates
        Int4ChooseQParamsAlgorithm(self.int4_choose_qparams_algorithm)   # validates


def _int4_weight_only_quantize_tensor(weight, config: Int4WeightOnlyConfig):
    group_size = config.group_size
    if weight.shape[-1] % group_size != 0:
        # reference :549-553: the layer is left unquantized
        logger.info(f"Skipping quantizing weight with int4 weight only quantization, because the shape of weight "
                    f"{weight.shape} is not compatible with group_size {group_size}")
        return weight
    block_size = list([1 for _ in range(weight.ndim - 1)] + [group_size])
    fmt = Int4PackingFormat(config.int4_packing_format)
    if fmt == Int4PackingFormat.TILE_PACKED_TO_4D:
        assert config.int4_tile_packed_ntile == 8, "ntile 16 is the ROCm variant; CUDA uses 8"
        return Int4TilePackedTo4dTensor.from_hp(weight, block_size,
                                                int4_choose_qparams_algorithm=Int4ChooseQParamsAlgorithm(
                                                    config.int4_choose_qparams_algorithm),
                                                ntile_size=config.int4_tile_packed_ntile)
    raise NotImplementedError(
        f"int4_packing_format={fmt.value!r}: only 'tile_packed_to_4d' has sm_100a kernels in this engine "
        f"(plain/preshuffled need the external mslk library in the reference as well)")


@register_quantize_module_handler(Int4WeightOnlyConfig)
def _int4_weight_only_transform(module: torch.nn.Module, config: Int4WeightOnlyConfig, *,
                                parameter_name: str = "weight") -> torch.nn.Module:
    assert hasattr(module, parameter_name), f"Expected module to have {parameter_name!r}"
    new_weight = _int4_weight_only_quantize_tensor(getattr(module, parameter_name), config)
    return _set_quantized_param(module, parameter_name, new_weight)


# ---------------------------------------------------------------------------------------------
# int8 dynamic activation x int8 weight
# ---------------------------------------------------------------------------------------------
def _validate_granularity_int8(act_granularity, weight_granularity):
    for g in (act_granularity, weight_granularity):
        if not isinstance(g, (PerRow, PerTensor)):
            raise ValueError(f"Unsupported granularity {g}: only PerTensor and PerRow are supported for int8")
        if isinstance(g, PerRow) and g.dim != -1:
            raise ValueError(f"Only PerRow(dim=-1) is supported, got {g}")


@dataclass
class Int8DynamicActivationInt8WeightConfig(AOBaseConfig):
    act_mapping_type: Optional[MappingType] = MappingType.SYMMETRIC
    weight_only_decode: bool = False
    granularity: Optional[Union[Granularity, List[Granularity]]] = PerRow()
    set_inductor_config: bool = True
    version: int = 2
    reduce_range: Optional[bool] = False

    def __post_init__(self):
        if self.version == 1:
            raise ValueError("version 1 of Int8DynamicActivationInt8WeightConfig has been removed, please use version 2")
        a, w = Int8Tensor._normalize_granularity(self.granularity)
        _validate_granularity_int8(a, w)
        assert self.act_mapping_type in (MappingType.SYMMETRIC, MappingType.ASYMMETRIC), (
            "Int8DynamicActivationInt8WeightConfig requires `act_mapping_type` in (MappingType.SYMMETRIC, "
            "MappingType.ASYMMETRIC).")


def _int8_dynamic_activation_int8_weight_quantize_tensor(weight, config):
    a, w = Int8Tensor._normalize_granularity(config.granularity)
    if weight.shape[-1] % 16 != 0 or weight.shape[-2] % 8 != 0:
        # the reference's cuBLAS path needs K % 8 == 0 and N % 8 == 0 and otherwise falls back to a CPU matmul
        # (int8/kernels.py:48-58); this engine has no fallback and its TMA row pitch needs K % 16 == 0: leave the
        # layer unquantized, with a log line, like the int4 / float8 flows do for incompatible shapes
        logger.info(f"Skipping int8 dynamic quantization: weight shape {tuple(weight.shape)} needs in_features % 16 == 0 "
                    f"and out_features % 8 == 0 for the sm_100a kernel")
        return weight
    return Int8Tensor.from_hp(
        weight, granularity=w, mapping_type=MappingType.SYMMETRIC,
        act_quant_kwargs=QuantizeTensorToInt8Kwargs(granularity=a, mapping_type=config.act_mapping_type,
                                                    reduce_range=bool(config.reduce_range)),
        reduce_range=config.reduce_range)


@register_quantize_module_handler(Int8DynamicActivationInt8WeightConfig)
def _int8_dynamic_activation_int8_weight_transform(module, config, *, parameter_name="weight"):
    assert hasattr(module, parameter_name), f"Expected module to have {parameter_name!r}"
    new_weight = _int8_dynamic_activation_int8_weight_quantize_tensor(getattr(module, parameter_name), config)
    return _set_quantized_param(module, parameter_name, new_weight)


# ---------------------------------------------------------------------------------------------
# float8 dynamic activation x float8 weight
# ---------------------------------------------------------------------------------------------
e4m3_dtype = torch.float8_e4m3fn


@dataclass
class Float8DynamicActivationFloat8WeightConfig(AOBaseConfig):
    activation_dtype: torch.dtype = e4m3_dtype
    weight_dtype: torch.dtype = e4m3_dtype
    granularity: Optional[Union[FP8Granularity, List[FP8Granularity]]] = None
    packing_format: Optional[Float8PackingFormat] = Float8PackingFormat.PLAIN
    mm_config: Optional[Float8MMConfig] = None
    activation_value_lb: Optional[float] = None
    activation_value_ub: Optional[float] = None
    kernel_preference: KernelPreference = KernelPreference.AUTO
    set_inductor_config: bool = True
    version: int = 2
    alg_id: int = 0

    def __post_init__(self):
        a, w = _normalize_granularity(self.granularity)
        self.granularity = [a, w]
        if self.mm_config is None:
            self.mm_config = Float8MMConfig(use_fast_accum=True)


def _fp8_mm_compat(weight: torch.Tensor) -> bool:
    """Both dims must be multiples of 16, else the layer is skipped (reference quantization/utils.py:663-687)."""
    assert weight.dim() in (2, 3), f"float8 quantization only works for 2/3-D tensors, got {weight.dim()}D"
    out_dim, in_dim = weight.shape[-2:]
    if in_dim % 16 != 0 or out_dim % 16 != 0:
        logger.info(f"Skipping float8 quantization: weight shape {weight.shape} is not compatible with _scaled_mm "
                    f"(both dims must be multiples of 16).")
        return False
    return True


def _float8_dynamic_activation_float8_weight_quantize_tensor(weight, config):
    if config.packing_format != Float8PackingFormat.PLAIN:
        raise NotImplementedError(f"float8 packing_format={config.packing_format}: only PLAIN is implemented")
    a_gran, w_gran = config.granularity
    if not _fp8_mm_compat(weight):
        return weight
    if isinstance(w_gran, PerRow):
        assert weight.dtype == torch.bfloat16, "PerRow quantization only works for bfloat16 precision input weight"
    act_quant_kwargs = QuantizeTensorToFloat8Kwargs(config.activation_dtype, a_gran, hp_value_lb=config.activation_value_lb,
                                                    hp_value_ub=config.activation_value_ub,
                                                    kernel_preference=config.kernel_preference)
    return Float8Tensor.from_hp(weight, float8_dtype=config.weight_dtype, granularity=w_gran, mm_config=config.mm_config,
                                kernel_preference=config.kernel_preference, act_quant_kwargs=act_quant_kwargs)


@register_quantize_module_handler(Float8DynamicActivationFloat8WeightConfig)
def _float8_dynamic_activation_float8_weight_transform(module, config, *, parameter_name="weight"):
    if torch.cuda.is_available():
        require_sm100()
    assert hasattr(module, parameter_name), f"Expected module to have {parameter_name!r}"
    new_weight = _float8_dynamic_activation_float8_weight_quantize_tensor(getattr(module, parameter_name), config)
    return _set_quantized_param(module, parameter_name, new_weight)


# ---------------------------------------------------------------------------------------------
# FqnToConfig (per-module / per-parameter configs; used by HF TorchAoConfig)
# ---------------------------------------------------------------------------------------------
@dataclass
class FqnToConfig(AOBaseConfig):
    """Ordered map key -> config (or None). Key: exact fqn of a module/parameter, ``re:<regex>``, or ``_default``."""

    fqn_to_config: "OrderedDict[str, Optional[AOBaseConfig]]" = field(default_factory=OrderedDict)
    module_fqn_to_config: "OrderedDict[str, Optional[AOBaseConfig]]" = field(default_factory=OrderedDict)
    version: int = 1

    def __post_init__(self):
        if len(self.module_fqn_to_config) > 0 and len(self.fqn_to_config) > 0 and \
                dict(self.module_fqn_to_config) != dict(self.fqn_to_config):
            raise ValueError("`fqn_to_config` and `module_fqn_to_config` are both specified and are not equal!")
        if len(self.module_fqn_to_config) > 0:
            self.fqn_to_config = self.module_fqn_to_config
        if len(self.fqn_to_config) > 0:
            self.module_fqn_to_config = self.fqn_to_config
        if any(k.startswith("re:") for k in self.fqn_to_config) and "_default" in self.fqn_to_config:
            logger.warning("`_default` with regex keys: regexes are tried first, `_default` is the fallback")


ModuleFqnToConfig = FqnToConfig


def fqn_matches_fqn_config(fqn: str, config: FqnToConfig) -> bool:
    if fqn in config.fqn_to_config:
        assert not fqn.startswith("re:"), f"Error: Exact match but regex {fqn} specified."
        return True
    return any(p.startswith("re:") and re.fullmatch(p[3:], fqn) for p in config.fqn_to_config)


def _top_level_params(module, fqn):
    for name, param in module.named_parameters():
        if name in dir(module):
            yield name, param, (f"{fqn}.{name}" if fqn else name)


def _module_param_matches_fqn_config(module, fqn, config) -> bool:
    return any(fqn_matches_fqn_config(pfqn, config) for _, _, pfqn in _top_level_params(module, fqn))


def _apply(module, c, parameter_name=None):
    if c is None:
        return module
    handler = _QUANTIZE_CONFIG_HANDLER[type(c)]
    return handler(module, c) if parameter_name is None else handler(module, c, parameter_name=parameter_name)


def _fqn_to_config_handler(module: torch.nn.Module, fqn: str, config: FqnToConfig):
    """Same order as the reference (quant_api.py:1638-1701): exact parameter fqns first (several may match; a `None`
    config takes the parameter out of the regex pass), then -- only when no parameter matched -- the exact module
    fqn; then EVERY regex that fully matches a remaining top-level parameter, in dict order; then -- only when still
    nothing matched -- the first module-fqn regex, and finally `_default`."""
    found = False
    remaining = []
    for name, _, pfqn in _top_level_params(module, fqn):
        if pfqn in config.fqn_to_config:
            found = True
            c = config.fqn_to_config[pfqn]
            if c is not None:
                module = _apply(module, c, parameter_name=name)
        else:
            remaining.append((name, pfqn))
    if not found and fqn in config.fqn_to_config:
        return _apply(module, config.fqn_to_config[fqn])
    for name, pfqn in remaining:
        for pat, c in config.fqn_to_config.items():
            if pat.startswith("re:") and re.fullmatch(pat[3:], pfqn):
                found = True
                module = _apply(module, c, parameter_name=name)
    if found:
        return module
    for pat, c in config.fqn_to_config.items():
        if pat.startswith("re:") and re.fullmatch(pat[3:], fqn):
            return _apply(module, c)
    if "_default" in config.fqn_to_config and _is_linear(module):
        return _apply(module, config.fqn_to_config["_default"])
    return module
