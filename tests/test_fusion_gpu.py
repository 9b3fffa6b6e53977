"""GPU tests of ao_b200.fusion (q|k|v and gate|up as one launch) and of row-strided activations.

Parity: the fused launch must reproduce the separate launches to fp32 re-association (stream-K splits a tile's K
range at different chunks for a different grid, so the fp32 partial sums associate differently: <= 1 bf16 ulp per
element, SQNR >= 60 dB), bit-exactly for int8 (integer accumulation + the reference rounding order), and the
member views must write through to the fused storage (narrow + copy_ loaders).
"""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _sqnr(ref, got):
    ref, got = ref.float(), got.float()
    d = (ref - got).norm()
    return float("inf") if d == 0 else float(20 * torch.log10(ref.norm() / d))


class Attn(nn.Module):
    def __init__(self, h=1024, kv=256, inter=2048, bias=False):
        super().__init__()
        mk = lambda k, n: nn.Linear(k, n, bias=bias, device="cuda", dtype=torch.bfloat16)
        self.q_proj, self.k_proj, self.v_proj = mk(h, h), mk(h, kv), mk(h, kv)
        self.gate_proj, self.up_proj = mk(h, inter), mk(h, inter)

    def forward(self, x):
        return self.q_proj(x), self.k_proj(x), self.v_proj(x), self.gate_proj(x), self.up_proj(x)


def _configs():
    from ao_b200.prototype.mx_formats import MXDynamicActivationMXWeightConfig, NVFP4WeightOnlyConfig
    from ao_b200.prototype.mx_formats.inference_workflow import NVFP4WeightFloat8ActivationConfig
    from ao_b200.quantization import (Float8DynamicActivationFloat8WeightConfig, Int4WeightOnlyConfig,
                                      Int8DynamicActivationInt8WeightConfig, PerRow)

    return {
        "int4": Int4WeightOnlyConfig(group_size=32, int4_packing_format="tile_packed_to_4d"),
        "int8": Int8DynamicActivationInt8WeightConfig(),
        "fp8": Float8DynamicActivationFloat8WeightConfig(granularity=PerRow()),
        "mxfp8": MXDynamicActivationMXWeightConfig(),
        "nvfp4w": NVFP4WeightOnlyConfig(),   # two-level scaling: each member keeps its own per-tensor scale (per out-feature)
        "nvfp4w_fp8a": NVFP4WeightFloat8ActivationConfig(),
    }


@pytest.mark.parametrize("fmt", ["int4", "int8", "fp8", "mxfp8", "nvfp4w", "nvfp4w_fp8a"])
@pytest.mark.parametrize("M,bias", [(1, False), (32, True), (5, False)])
def test_fused_matches_separate(fmt, M, bias):
    import ao_b200  # noqa: F401
    from ao_b200.fusion import FusedLinearMember, fuse_parallel_linears
    from ao_b200.quantization import quantize_

    torch.manual_seed(0)
    m = Attn(bias=bias)
    quantize_(m, _configs()[fmt])
    x = torch.randn(M, 1024, device="cuda", dtype=torch.bfloat16)
    with torch.no_grad():
        ref = m(x)
        n = torch.ops.ao_b200.launch_count()
        ref = m(x)
        launches_sep = torch.ops.ao_b200.launch_count() - n
    assert fuse_parallel_linears(m) == 2
    assert all(isinstance(getattr(m, k), FusedLinearMember) for k in ("q_proj", "k_proj", "v_proj", "gate_proj", "up_proj"))
    with torch.no_grad():
        got = m(x)
        n = torch.ops.ao_b200.launch_count()
        got = m(x)
        launches_fused = torch.ops.ao_b200.launch_count() - n
    assert launches_fused < launches_sep
    for r, g in zip(ref, got):
        assert r.shape == g.shape
        if fmt == "int8":
            assert torch.equal(r, g)   # integer accumulation: exact whatever the split
        else:
            assert _sqnr(r, g) > 60.0
    # the same buffer with NEW contents (a CUDA graph's static input): nothing stale may be served
    x.copy_(torch.randn_like(x))
    torch.cuda.synchronize()
    with torch.no_grad():
        again = m(x)
        want = tuple(F.linear(x, getattr(m, n)._group.weight, getattr(m, n)._group.bias) for n in ("q_proj", "gate_proj"))
    assert torch.equal(again[0], want[0][..., : again[0].shape[-1]]) and torch.equal(again[3], want[1][..., : again[3].shape[-1]])
    # a member called with a different input on its own still answers for that input
    x2 = torch.randn(M, 1024, device="cuda", dtype=torch.bfloat16)
    with torch.no_grad():
        k2 = m.k_proj(x2)
        k_ref = m(x2)[1]
    assert torch.equal(k2, k_ref)


def test_member_weights_are_views_of_the_fused_storage():
    import ao_b200  # noqa: F401
    from ao_b200.fusion import fuse_parallel_linears
    from ao_b200.quantization import quantize_

    torch.manual_seed(1)
    m, donor = Attn(), Attn()
    quantize_(m, _configs()["int4"])
    quantize_(donor, _configs()["int4"])
    fuse_parallel_linears(m)
    x = torch.randn(4, 1024, device="cuda", dtype=torch.bfloat16)
    with torch.no_grad():
        want = donor.k_proj(x)
        # a TP-style loader: narrow the parameter, copy_ the loaded shard (reference: torchao/testing/utils.py:496-519)
        w = m.k_proj.weight.data
        w.narrow(0, 0, w.shape[0]).copy_(donor.k_proj.weight.data)
        got = m(x)[1]
    assert torch.equal(got, want)


@pytest.mark.parametrize("M", [1, 7, 32, 100])
def test_row_strided_activation_needs_no_copy(M):
    """x = a column slice of a wider buffer: same result as its contiguous copy, bit for bit."""
    import ao_b200  # noqa: F401

    ops = torch.ops.ao_b200
    gen = torch.Generator(device="cuda").manual_seed(3)
    N, K, g = 512, 2048, 32
    qd = torch.randint(-2**31, 2**31 - 1, (N // 8, K // 128, 32, 4), device="cuda", dtype=torch.int32, generator=gen)
    sz = ((torch.rand(K // g, N, 2, device="cuda", generator=gen) - 0.5) * 0.02).to(torch.bfloat16)
    wide = torch.randn(M, K + 1024, device="cuda", generator=gen).to(torch.bfloat16)
    for off in (0, 1024):
        xs = wide[:, off: off + K]
        assert not xs.is_contiguous() or M == 1
        for impl in (1, 2):
            y_s = ops.int4_tilepacked_linear(xs, qd, g, sz, None, N, impl)
            y_c = ops.int4_tilepacked_linear(xs.contiguous(), qd, g, sz, None, N, impl)
            assert torch.equal(y_s, y_c)


def test_llama_layer_chain_fused_vs_unfused():
    """The bench model: fused (4 launches / layer, strided slices between them) vs unfused (7 launches / layer)."""
    import ao_b200  # noqa: F401
    from ao_b200.fusion import fuse_parallel_linears
    from ao_b200.models import LlamaLinearStack, LlamaShape
    from ao_b200.quantization import Int4WeightOnlyConfig, quantize_

    shape = LlamaShape("tiny", 1024, 3072, 256, 2)
    a = LlamaLinearStack(shape, device="cuda", seed=0, init_scale=0.05)
    b = LlamaLinearStack(shape, device="cuda", seed=0, init_scale=0.05)
    cfg = Int4WeightOnlyConfig(group_size=32, int4_packing_format="tile_packed_to_4d")
    quantize_(a, cfg)
    quantize_(b, cfg)
    assert fuse_parallel_linears(b) == 4
    for M in (1, 32):
        x = torch.randn(M, 1024, device="cuda", dtype=torch.bfloat16)
        with torch.no_grad():
            n0 = torch.ops.ao_b200.launch_count()
            ya = a(x)
            n1 = torch.ops.ao_b200.launch_count()
            yb = b(x)
            n2 = torch.ops.ao_b200.launch_count()
        assert (n1 - n0, n2 - n1) == (14, 8)
        with torch.no_grad():   # a second pass over the same buffer launches everything again (graph capture relies on it)
            b(x)
        assert torch.ops.ao_b200.launch_count() - n2 == 8
        assert _sqnr(ya, yb) > 40.0   # two layers of bf16 re-rounding between differently associated sums


@pytest.mark.parametrize("M", [1, 5, 32, 130])
def test_row_strided_quantizers_match_contiguous(M):
    """The activation quantizers take a column slice of a wider buffer directly: same bytes as for its contiguous copy
    (data, scales -- including the zero padding of the blocked scale layouts, which the kernels now write themselves)."""
    import ao_b200  # noqa: F401

    ops = torch.ops.ao_b200
    gen = torch.Generator(device="cuda").manual_seed(5)
    K = 1056   # not a multiple of 128: the blocked scale layouts have padding columns
    wide = torch.randn(M, K + 512, device="cuda", generator=gen).to(torch.bfloat16)
    xs = wide[:, 256: 256 + K]
    xc = xs.contiguous()
    pts = torch.tensor([0.01], device="cuda")
    for fn in (lambda t: ops.int8_quantize_rowwise(t), lambda t: ops.fp8_quantize_rowwise(t),
               lambda t: ops.mxfp8_quantize(t, True), lambda t: ops.mxfp8_quantize(t, False),
               lambda t: ops.nvfp4_quantize(t, pts, True), lambda t: ops.nvfp4_quantize(t, None, False),
               lambda t: ops.fp8_fakequant_rowwise(t)):
        a, b = fn(xs), fn(xc)
        for u, v in zip(a, b):
            assert torch.equal(u.view(torch.uint8) if u.element_size() == 1 else u, v.view(torch.uint8) if v.element_size() == 1 else v)
    # padding entries of the blocked layouts are zero (they multiply TMA's zero fill in the GEMM: must not be NaN)
    from oracle import oracle as o

    _, s = ops.mxfp8_quantize(xs, True)
    ref_q, ref_s = o.mxfp8_quantize(o.bf16_bits(xc))
    assert np.array_equal(s.cpu().numpy().reshape(-1), o.to_blocked(ref_s).reshape(-1))
