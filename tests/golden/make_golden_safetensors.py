"""Golden fixture for the safetensors wire format: flatten a state dict holding one tensor of each hot-path class with
the REFERENCE (torchao 0.19, CPU) and store the metadata strings plus the flat tensors.

Run in the build container only (needs /root/reference):
    PYTHONPATH=/root/reference python tests/golden/make_golden_safetensors.py
Writes safetensors_wire.json (metadata, key -> dtype/shape) and safetensors_wire.npz (tensor bytes as uint8).
"""
import json
import os
import sys

import numpy as np
import torch

REF = os.environ.get("AO_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    torch.manual_seed(7)
    from torchao.prototype.mx_formats.mx_tensor import MXTensor, QuantizeTensorToMXKwargs
    from torchao.prototype.mx_formats.nvfp4_tensor import NVFP4Tensor, QuantizeTensorToNVFP4Kwargs
    from torchao.prototype.safetensors.safetensors_support import flatten_tensor_state_dict
    from torchao.quantization import Float8Tensor, Int4TilePackedTo4dTensor, Int8Tensor, PerRow
    from torchao.quantization.quantize_.workflows import QuantizeTensorToFloat8Kwargs, QuantizeTensorToInt8Kwargs

    N, K = 32, 128
    w = (torch.randn(N, K) * 0.05).to(torch.bfloat16)
    sd = {}
    sd["m.int4.weight"] = Int4TilePackedTo4dTensor(
        torch.randint(-2**31, 2**31 - 1, (N // 8, K // 128, 32, 4), dtype=torch.int32),
        (torch.rand(K // 32, N, 2) * 0.01).to(torch.bfloat16), [1, 32], [N, K])
    sd["m.int8.weight"] = Int8Tensor.from_hp(w, granularity=PerRow(),
                                             act_quant_kwargs=QuantizeTensorToInt8Kwargs(granularity=PerRow()))
    sd["m.fp8.weight"] = Float8Tensor.from_hp(w, granularity=PerRow(),
                                              act_quant_kwargs=QuantizeTensorToFloat8Kwargs(granularity=PerRow()))
    sd["m.mx.weight"] = MXTensor.to_mx(w, torch.float8_e4m3fn, 32,
                                       act_quant_kwargs=QuantizeTensorToMXKwargs(elem_dtype=torch.float8_e4m3fn, block_size=32))
    sd["m.nvfp4.weight"] = NVFP4Tensor.to_nvfp4(w, per_tensor_scale=(w.float().abs().max() / (448.0 * 6.0)),
                                                act_quant_kwargs=QuantizeTensorToNVFP4Kwargs())
    sd["m.int8.bias"] = torch.randn(N).to(torch.bfloat16)
    flat, meta = flatten_tensor_state_dict(sd)
    index = {k: {"dtype": str(v.dtype).split(".")[-1], "shape": list(v.shape)} for k, v in flat.items()}
    arrays = {k.replace(".", "__"): v.detach().contiguous().reshape(-1).view(torch.uint8).numpy().copy() for k, v in flat.items()}
    with open(os.path.join(HERE, "safetensors_wire.json"), "w") as f:
        json.dump({"metadata": meta, "tensors": index}, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "safetensors_wire.npz"), **arrays)
    print("wrote", len(flat), "tensors;", {k: json.loads(v).get("_type") if k != "tensor_names" else "-" for k, v in meta.items()})


if __name__ == "__main__":
    main()
