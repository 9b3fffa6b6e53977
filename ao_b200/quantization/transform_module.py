"""Registry config-type -> module transform (reference: torchao/quantization/transform_module.py:19-52)."""
import functools
from typing import Callable, Dict, Type

import torch

from ao_b200.core.config import AOBaseConfig

_QUANTIZE_CONFIG_HANDLER: Dict[Type[AOBaseConfig], Callable[..., torch.nn.Module]] = {}


def register_quantize_module_handler(config_type):
    """``@register_quantize_module_handler(Cfg)`` on ``fn(module, config, *, parameter_name="weight")``."""

    @functools.wraps(config_type)
    def decorator(func):
        _QUANTIZE_CONFIG_HANDLER[config_type] = func
        return func

    return decorator
