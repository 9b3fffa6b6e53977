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


def test_int4_narrow_copy_writes_through_to_the_parameter():
    """vLLM's loader (reference torchao/testing/utils.py:471-519): narrow both sides, copy_ in place.  The narrowed
    tensor must alias the parameter's storage for BOTH payload tensors, otherwise the loaded shard is lost."""
    import ao_b200  # noqa: F401
    from ao_b200.quantization import Int4TilePackedTo4dTensor

    N, K, g = 1024, 1024, 32

    def make(seed):
        gen = torch.Generator().manual_seed(seed)
        return Int4TilePackedTo4dTensor(torch.randint(-2**31, 2**31 - 1, (N // 8, K // 128, 32, 4), dtype=torch.int32, generator=gen),
                                        torch.rand(K // g, N, 2, generator=gen).to(torch.bfloat16), [1, g], [N, K])

    param, loaded = make(0), make(1)
    for rank in (0, 1):
        dst = param.narrow(0, rank * 512, 512)
        src = loaded.narrow(0, rank * 512, 512)
        assert dst.qdata.data_ptr() == param.qdata[rank * 64:].data_ptr()      # aliasing (reference test_slice_preserves_aliasing)
        assert not torch.equal(dst.qdata[0], src.qdata[0])
        dst.copy_(src)
        assert torch.equal(dst.qdata, src.qdata) and torch.equal(dst.scale_and_zero, src.scale_and_zero)
    assert torch.equal(param.qdata, loaded.qdata) and torch.equal(param.scale_and_zero, loaded.scale_and_zero)
    # K-dim narrow (row-parallel layers): multiples of 1024 only
    half = make(2)
    big = Int4TilePackedTo4dTensor(torch.zeros(N // 8, 2 * K // 128, 32, 4, dtype=torch.int32),
                                   torch.zeros(2 * K // g, N, 2, dtype=torch.bfloat16), [1, g], [N, 2 * K])
    big.narrow(1, K, K).copy_(half)
    assert torch.equal(big.qdata[:, K // 128:], half.qdata) and torch.equal(big.scale_and_zero[K // g:], half.scale_and_zero)
    assert int(big.qdata[:, : K // 128].abs().sum()) == 0


@pytest.mark.parametrize("kind", ["int8", "fp8"])
def test_rowwise_classes_narrow_copy_aliasing(kind):
    """reference test_float8_tensor.py::test_slice_preserves_aliasing / test_slice_and_copy_similar_to_vllm and the int8
    equivalents: narrow returns views of qdata and of the per-row scale; copy_ through them updates the parameter."""
    import ao_b200  # noqa: F401
    from ao_b200.quantization import Float8Tensor, Int8Tensor

    def make(seed):
        gen = torch.Generator().manual_seed(seed)
        if kind == "int8":
            return Int8Tensor(torch.randint(-128, 127, (1024, 512), dtype=torch.int8, generator=gen),
                              torch.rand(1024, 1, generator=gen), [1, 512], torch.bfloat16)
        return Float8Tensor(torch.randn(1024, 512, generator=gen).to(torch.float8_e4m3fn), torch.rand(1024, 1, generator=gen), [1, 512])

    param, loaded = make(0), make(1)
    view = param.narrow(0, 0, 512)
    assert view.qdata.data_ptr() == param.qdata.data_ptr() and view.scale.data_ptr() == param.scale.data_ptr()
    for rank in (0, 1):
        dst, src = param.narrow(0, rank * 512, 512), loaded.narrow(0, rank * 512, 512)
        dst.copy_(src)
        assert torch.equal(dst.qdata.view(torch.uint8), src.qdata.view(torch.uint8)) and torch.equal(dst.scale, src.scale)
    assert torch.equal(param.qdata.view(torch.uint8), loaded.qdata.view(torch.uint8)) and torch.equal(param.scale, loaded.scale)
    cols = param.narrow(1, 0, 256)      # K-dim narrow of a rowwise tensor keeps the whole scale column
    assert cols.qdata.shape == (1024, 256) and cols.scale.shape == (1024, 1)
    assert cols.scale.data_ptr() == param.scale.data_ptr()
