"""ao_b200 — B200-native quantized-linear engine with torchao's operator surface.

Drop-in for the quantized ``nn.Linear`` forward of pytorch/ao (torchao 0.19): the same
``quantize_`` / config / tensor-subclass API, with the kernels underneath replaced by
hand-written sm_100a CUDA (``libao_b200.so``, C ABI in ``include/ao_b200.h``) registered as
``torch.ops.ao_b200.*``.  There is no CPU or eager fallback for the hot path: importing this
package without the native library raises.
"""
from ._native import load_native, native_lib_paths  # noqa: F401

load_native()

__version__ = "0.1.0"

from . import quantization  # noqa: E402,F401
from .quantization import quantize_  # noqa: E402,F401
