"""aten.slice / narrow / copy_ of the quantized tensor classes, as tensor-parallel weight loaders use them
(`param.data.narrow(dim, start, size).copy_(loaded.narrow(...))`, reference torchao/testing/utils.py:471-519).
NVFP4 / MX payloads are checked byte for byte against slices taken by the reference (tests/golden/make_golden_slices.py)."""
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SLICES = {"r0": (0, 0, 128), "r1": (0, 128, 256), "c0": (1, 0, 128), "c1": (1, 64, 192), "c2": (1, 128, 256)}


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "mx_nvfp4_slices.npz"))


def _u8(t):
    return t.detach().contiguous().reshape(-1).view(torch.uint8).numpy()


def _build(gold, case):
    import ao_b200  # noqa: F401
    from ao_b200.prototype.mx_formats import MXTensor, NVFP4Tensor
    from ao_b200.quantization.quantize_.common.kernel_preference import KernelPreference

    M, K = 256, 256
    sshape = tuple(int(v) for v in gold[f"{case}__s_shape"])
    q = torch.from_numpy(gold[f"{case}__q"].copy())
    s = torch.from_numpy(gold[f"{case}__s"].copy())
    if case.startswith("nv"):
        return NVFP4Tensor(q.reshape(M, K // 2), s.view(torch.float8_e4m3fn).reshape(sshape), 16, torch.bfloat16,
                           torch.tensor(0.01), None, case == "nv_blocked", False, None)
    return MXTensor(q.view(torch.float8_e4m3fn).reshape(M, K), s.view(torch.float8_e8m0fnu).reshape(sshape),
                    torch.float8_e4m3fn, 32, torch.bfloat16, KernelPreference.AUTO, None, False)


@pytest.mark.parametrize("case", ["nv_plain", "nv_blocked", "mx_plain"])
@pytest.mark.parametrize("sl", list(SLICES))
def test_slices_match_reference_bytes(gold, case, sl):
    t = _build(gold, case)
    dim, a, b = SLICES[sl]
    if case == "nv_blocked" and sl == "c1":
        pass  # 64..192: both ends are multiples of 64, allowed
    s = t.narrow(dim, a, b - a)
    shapes = [int(v) for v in gold[f"{case}__{sl}__shapes"]]
    assert list(s.shape) + list(s.qdata.shape) + list(s.scale.shape) == shapes
    assert np.array_equal(_u8(s.qdata), gold[f"{case}__{sl}__q"])
    assert np.array_equal(_u8(s.scale), gold[f"{case}__{sl}__s"])
    assert type(s) is type(t) and s.is_swizzled_scales == t.is_swizzled_scales


def test_blocked_scale_alignment_errors(gold):
    t = _build(gold, "nv_blocked")
    with pytest.raises(RuntimeError):
        t.narrow(0, 64, 128)      # rows of a blocked scale layout move in tiles of 128
    with pytest.raises(RuntimeError):
        t.narrow(1, 32, 64)       # columns in tiles of 4 scales = 64 elements
    with pytest.raises(ValueError):
        t[::2]
    p = _build(gold, "nv_plain")
    assert p[3].shape == (256,) and p[3].qdata.shape == (128,)
    with pytest.raises(AssertionError):
        t[3]                      # select on blocked scales is unsupported, as in the reference


def test_narrow_copy_loader_pattern(gold):
    """Fill a full-size parameter shard by shard, the way a tensor-parallel loader does."""
    full = _build(gold, "nv_plain")
    dst = _build(gold, "nv_plain")
    dst.qdata.zero_()
    dst.scale.view(torch.uint8).zero_()
    for a in (0, 128):
        dst.narrow(0, a, 128).copy_(full.narrow(0, a, 128))
    assert torch.equal(dst.qdata, full.qdata) and torch.equal(dst.scale.view(torch.uint8), full.scale.view(torch.uint8))

    from ao_b200.quantization import Float8Tensor, Int8Tensor  # rowwise classes slice on both dims
    q = torch.randint(-128, 127, (64, 128), dtype=torch.int8)
    t8 = Int8Tensor(q, torch.rand(64, 1), [1, 128], torch.bfloat16)
    assert torch.equal(t8.narrow(0, 16, 32).qdata, q[16:48]) and t8.narrow(0, 16, 32).scale.shape == (32, 1)
    assert torch.equal(t8.narrow(1, 0, 64).qdata, q[:, :64]) and t8.narrow(1, 0, 64).scale.shape == (64, 1)
    f8 = Float8Tensor(q.view(torch.float8_e4m3fn) if False else torch.zeros(64, 128).to(torch.float8_e4m3fn), torch.rand(64, 1), [1, 128])
    assert f8.narrow(0, 0, 16).scale.shape == (16, 1)
