"""Golden fixture for aten.slice of NVFP4Tensor / MXTensor (plain and blocked scales), from the REFERENCE on CPU.
    PYTHONPATH=/root/reference python tests/golden/make_golden_slices.py
Writes mx_nvfp4_slices.npz: inputs (bf16 bits), the full tensors' payloads and the payloads of each slice."""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("AO_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))


def u8(t):
    return t.detach().contiguous().reshape(-1).view(torch.uint8).numpy().copy()


def main():
    torch.manual_seed(11)
    from torchao.prototype.mx_formats.mx_tensor import MXTensor
    from torchao.prototype.mx_formats.nvfp4_tensor import NVFP4Tensor

    M, K = 256, 256
    w = (torch.randn(M, K) * 0.05).to(torch.bfloat16)
    out = {"w_bits": w.view(torch.int16).numpy().view(np.uint16).copy()}
    pts = w.float().abs().max() / (448.0 * 6.0)
    cases = {
        "nv_plain": NVFP4Tensor.to_nvfp4(w, per_tensor_scale=pts, is_swizzled_scales=False),
        "nv_blocked": NVFP4Tensor.to_nvfp4(w, per_tensor_scale=pts, is_swizzled_scales=True),
        "mx_plain": MXTensor.to_mx(w, torch.float8_e4m3fn, 32),
    }
    slices = {"r0": (0, 0, 128), "r1": (0, 128, 256), "c0": (1, 0, 128), "c1": (1, 64, 192), "c2": (1, 128, 256)}
    for cname, t in cases.items():
        out[f"{cname}__q"] = u8(t.qdata)
        out[f"{cname}__s"] = u8(t.scale)
        out[f"{cname}__s_shape"] = np.array(t.scale.shape)
        for sname, (dim, a, b) in slices.items():
            s = t.narrow(dim, a, b - a)
            out[f"{cname}__{sname}__q"] = u8(s.qdata)
            out[f"{cname}__{sname}__s"] = u8(s.scale)
            out[f"{cname}__{sname}__shapes"] = np.array(list(s.shape) + list(s.qdata.shape) + list(s.scale.shape))
    np.savez_compressed(os.path.join(HERE, "mx_nvfp4_slices.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
