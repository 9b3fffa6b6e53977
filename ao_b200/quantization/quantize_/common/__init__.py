from .kernel_preference import KernelPreference  # noqa: F401
from .quantize_tensor_kwargs import QuantizeTensorKwargs, _choose_quant_func_and_quantize_tensor  # noqa: F401
