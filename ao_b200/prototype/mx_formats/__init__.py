from .inference_workflow import (  # noqa: F401
    MXDynamicActivationMXWeightConfig, NVFP4DynamicActivationNVFP4WeightConfig, NVFP4WeightFloat8ActivationConfig,
    NVFP4WeightOnlyConfig, QuantizationStep)
from .mx_tensor import MXTensor, QuantizeTensorToMXKwargs, ScaleCalculationMode  # noqa: F401
from .nvfp4_tensor import (  # noqa: F401
    NVFP4Tensor, QuantizeTensorToFloat8ActKwargs, QuantizeTensorToNVFP4Kwargs, per_tensor_amax_to_scale)
