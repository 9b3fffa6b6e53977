"""Golden fixture for HQQ int4 qparams from the REFERENCE on CPU (fp32 solver path).
    PYTHONPATH=/root/reference python tests/golden/make_golden_hqq.py
Writes int4_hqq.npz: per case the bf16 weight bits, the reference's codes, scale and zero (bf16 bits)."""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("AO_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))


def bits(t):
    return t.detach().contiguous().view(torch.int16).numpy().view(np.uint16).copy()


def main():
    from torchao.quantization.quant_primitives import _choose_qparams_and_quantize_affine_hqq

    out = {}
    for name, (N, K, g, seed, scale) in {"g32": (64, 1024, 32, 5, 0.02), "g128": (32, 2048, 128, 6, 0.05),
                                        "g64_outlier": (16, 1024, 64, 7, 0.02)}.items():
        torch.manual_seed(seed)
        w = torch.randn(N, K) * scale
        if "outlier" in name:
            w[::3, ::97] *= 12.0
        w = w.to(torch.bfloat16)
        q, s, z, _ = _choose_qparams_and_quantize_affine_hqq(w, nbits=4, group_size=g, axis=1, compute_dtype=torch.bfloat16,
                                                             device="cpu", verbose=False, raw_output=False)
        out[f"{name}__w"] = bits(w)
        out[f"{name}__q"] = q.numpy().copy()
        out[f"{name}__s"] = bits(s).reshape(N, K // g)
        out[f"{name}__z"] = bits(z).reshape(N, K // g)
        out[f"{name}__g"] = np.array(g)
    np.savez_compressed(os.path.join(HERE, "int4_hqq.npz"), **out)
    print("wrote", sorted(out))


if __name__ == "__main__":
    main()
