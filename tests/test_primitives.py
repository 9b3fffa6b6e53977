"""CPU parity of the eager quantization primitives (ao_b200/quantization/quant_primitives.py, setup-time only)
against the oracle (itself pinned bit-exactly to reference fixtures by tests/test_oracle_golden.py).

Bit-exact: tinygemm qparams computed in bf16 arithmetic (reference quant_primitives.py:1268-1335 in eager), the 4-bit
codes, the packed (scale, zero) layout, per-row int8 scales + codes and per-row e4m3 scales + codes."""
import numpy as np
import pytest
import torch


@pytest.fixture(scope="module")
def o():
    from oracle import oracle as o

    return o


def _w(n, k, seed, scale=0.02):
    g = torch.Generator().manual_seed(seed)
    w = (torch.randn(n, k, generator=g) * scale).to(torch.bfloat16)
    w[0, :32] = 0        # an all-zero group (eps clamp of the scale)
    w[1, 32:64] = 0.5    # a constant group (max == min)
    return w


@pytest.mark.parametrize("g", [32, 64, 128, 256])
def test_tinygemm_qparams_and_codes_bit_exact(o, g):
    import ao_b200  # noqa: F401
    from ao_b200.quantization.quant_primitives import choose_qparams_affine_tinygemm, quantize_affine_tinygemm
    from ao_b200.quantization.utils import pack_tinygemm_scales_and_zeros

    w = _w(64, 1024, g)
    s, z = choose_qparams_affine_tinygemm(w, g)
    q = quantize_affine_tinygemm(w, g, s, z)
    s_o, z_o = o.int4_choose_qparams(o.bf16_bits(w), g)
    assert np.array_equal(o.bf16_bits(s), s_o) and np.array_equal(o.bf16_bits(z), z_o)
    assert np.array_equal(q.numpy().astype(np.uint8), o.int4_quantize(o.bf16_bits(w), g, s_o, z_o))
    assert np.array_equal(o.bf16_bits(pack_tinygemm_scales_and_zeros(s, z, s.dtype)), o.pack_scales_and_zeros(s_o, z_o))


def test_int8_rowwise_bit_exact(o):
    import ao_b200  # noqa: F401
    from ao_b200.quantization.quant_primitives import choose_qparams_affine_int8, quantize_affine_int8

    x = _w(48, 512, 3, scale=1.0)
    x[5] = 0   # all-zero row: eps clamp
    bs = [1, 512]
    s, zp = choose_qparams_affine_int8(x, bs)
    q = quantize_affine_int8(x, bs, s, zp)
    q_o, s_o = o.int8_quantize_rowwise(o.bf16_bits(x))
    assert np.array_equal(s.reshape(-1).float().numpy(), s_o.reshape(-1))
    assert np.array_equal(q.numpy(), q_o)


def test_fp8_rowwise_bit_exact(o):
    import ao_b200  # noqa: F401
    from ao_b200.quantization.quant_primitives import choose_scale_float8, quantize_affine_float8

    x = _w(48, 512, 4, scale=1.0)
    bs = [1, 512]
    s = choose_scale_float8(x, bs)
    q = quantize_affine_float8(x, s)
    q_o, s_o = o.fp8_quantize_rowwise(o.bf16_bits(x))
    assert np.array_equal(s.reshape(-1).float().numpy(), s_o.reshape(-1))
    assert np.array_equal(q.view(torch.uint8).numpy(), q_o)
