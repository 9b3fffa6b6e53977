"""Pins the CPU oracle (oracle/ao_oracle.c) to the reference's own arithmetic.

Fixtures under tests/golden/ were produced by importing torchao 0.19 from /root/reference on
CPU (tests/golden/make_golden.py); nothing here reads /root/reference.
Integer / byte / index results must be bit-exact.
"""
import os

import numpy as np
import pytest

from oracle import oracle as o


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("g", [32, 128])
def test_int4_tinygemm_qparams_and_quantize_bit_exact(golden_dir, g):
    d = load(golden_dir, "int4_tinygemm.npz")
    w = d[f"w_g{g}"]
    s, z = o.int4_choose_qparams(w, g)
    assert np.array_equal(s, d[f"s_g{g}"]), "scale differs from _choose_qparams_affine_tinygemm"
    assert np.array_equal(z, d[f"z_g{g}"]), "zero_point differs from _choose_qparams_affine_tinygemm"
    q = o.int4_quantize(w, g, s, z)
    assert np.array_equal(q, d[f"q_g{g}"]), "q differs from _quantize_affine_tinygemm"
    sz = o.pack_scales_and_zeros(s, z)
    assert np.array_equal(sz, d[f"sz_g{g}"]), "pack_tinygemm_scales_and_zeros layout differs"


def test_int4_pack_roundtrip_and_layout_formula():
    rng = np.random.default_rng(0)
    for (N, K, ikt) in [(8, 128, 8), (24, 1024, 8), (16, 256, 4), (16, 64, 2)]:
        q = rng.integers(0, 16, size=(N, K), dtype=np.uint8)
        qd = o.int4_pack_tile4d(q, ikt)
        assert qd.shape == (N // 8, K // (ikt * 16), 32, ikt // 2)
        assert np.array_equal(o.int4_unpack_tile4d(qd), q)
    # spot-check the documented index formula: word(n8,ko,t,wd) nibble e / half h
    N, K = 16, 256
    q = rng.integers(0, 16, size=(N, K), dtype=np.uint8)
    qd = o.int4_pack_tile4d(q, 8).view(np.uint32)
    for (n8, ko, t, wd) in [(0, 0, 0, 0), (1, 1, 31, 3), (0, 1, 13, 2), (1, 0, 6, 1)]:
        w = int(qd[n8, ko, t, wd])
        n = 8 * n8 + t // 4
        k0 = 128 * ko + 32 * wd + 2 * (t % 4)
        for e in range(4):
            assert (w >> (4 * e)) & 15 == q[n, k0 + 8 * e]
            assert (w >> (16 + 4 * e)) & 15 == q[n, k0 + 8 * e + 1]


def test_int4_cpu_linear_matches_reference_cpu_path(golden_dir):
    """BASELINE config[0]: the reference's Int4OpaqueTensor CPU linear (aten._weight_int4pack_mm_for_cpu)."""
    d = load(golden_dir, "int4_cpu_linear.npz")
    g = int(d["g"])
    w, x, bias, y_ref, sz_ref = d["w"], d["x"], d["bias"], d["y"], d["sz"]
    s, z = o.int4_choose_qparams(w, g)
    q = o.int4_quantize(w, g, s, z)
    sz = o.pack_scales_and_zeros(s, z)
    assert np.array_equal(sz, sz_ref), "scale_and_zero differs from the reference tensor's"
    qd = o.int4_pack_tile4d(q, 8) if w.shape[1] % 128 == 0 else None
    w_hat = o.int4_dequant(q, sz, g)
    y = o.linear_f32(o.bf16_to_f32(x), o.bf16_to_f32(w_hat), o.bf16_to_f32(bias))
    y_ref_f = o.bf16_to_f32(y_ref)
    # the CPU kernel accumulates in fp32 in its own order and rounds to bf16: >= 45 dB, i.e. bf16 rounding noise
    assert o.sqnr_db(y, y_ref_f) > 45.0
    if qd is not None:
        y2 = o.int4_linear(x, qd, sz, g, bias)
        assert o.sqnr_db(y_ref_f, o.bf16_to_f32(y2)) > 45.0


def test_int8_rowwise_quant_bit_exact(golden_dir):
    d = load(golden_dir, "int8_rowwise.npz")
    xq, xs = o.int8_quantize_rowwise(d["x"])
    assert np.array_equal(xs, d["xs"]) and np.array_equal(xq, d["xq"])
    wq, ws = o.int8_quantize_rowwise(d["w"])
    assert np.array_equal(ws, d["ws"]) and np.array_equal(wq, d["wq"])


def test_int8_linear_matches_reference_cpu(golden_dir):
    d = load(golden_dir, "int8_rowwise.npz")
    acc = o.int8_mm(d["xq"], d["wq"])
    # exact integer check against numpy
    assert np.array_equal(acc, d["xq"].astype(np.int32) @ d["wq"].astype(np.int32).T)
    y = o.int8_epilogue(acc, d["xs"], d["ws"], d["bias"])
    y_ref = d["y"]
    # the reference CPU path (int8/kernels.py:79-111) applies the same two-step scaling
    diff = np.abs(o.bf16_to_f32(y) - o.bf16_to_f32(y_ref))
    ulp = np.maximum(np.abs(o.bf16_to_f32(y_ref)), 1e-30) * 2.0 ** -7
    assert np.all(diff <= ulp + 1e-12), f"max diff {diff.max()}"
    assert (y != y_ref).mean() < 0.02


def test_fp8_codec_and_rowwise_bit_exact(golden_dir):
    d = load(golden_dir, "fp8_rowwise.npz")
    table = o.e4m3_to_f32(np.arange(256, dtype=np.uint8))
    ref = d["e4m3_table"]
    assert np.array_equal(np.isnan(table), np.isnan(ref))
    assert np.array_equal(table[~np.isnan(ref)], ref[~np.isnan(ref)])
    enc = np.array([o.lib().ao_oracle_f32_to_e4m3(float(min(max(v, -448.0), 448.0))) for v in d["sweep"]], np.uint8)
    assert np.array_equal(enc, d["sweep_q"])
    q, s = o.fp8_quantize_rowwise(d["x"])
    assert np.array_equal(s, d["s"])
    assert np.array_equal(q, d["q"])


def test_mxfp8_rceil_bit_exact(golden_dir):
    d = load(golden_dir, "mxfp8.npz")
    e = np.array([o.lib().ao_oracle_f32_to_e8m0_rceil(float(v)) for v in d["e8m0_in"]], np.uint8)
    assert np.array_equal(e, d["e8m0_out"])
    q, s = o.mxfp8_quantize(d["x"])
    assert np.array_equal(s, d["scale"])
    assert np.array_equal(q, d["q"])
    blocked = o.to_blocked(s)
    assert np.array_equal(blocked.reshape(-1), d["blocked_of_plain"].reshape(-1))
    assert np.array_equal(blocked.reshape(-1), d["scale_swizzled"].reshape(-1))
    assert np.array_equal(o.from_blocked(blocked, *s.shape), s)
    dq = o.f32_to_bf16(o.mxfp8_dequant(q, s))
    assert np.array_equal(dq, d["dq"])


def test_nvfp4_bit_exact(golden_dir):
    d = load(golden_dir, "nvfp4.npz")
    tab = np.array([o.lib().ao_oracle_e2m1_to_f32(i) for i in range(16)], np.float32)
    assert np.array_equal(tab, d["f4_table"])
    enc = np.array([o.lib().ao_oracle_f32_to_e2m1(float(v)) for v in d["f4_sweep"]], np.uint8)
    assert np.array_equal(enc, d["f4_sweep_q"])
    q1, s1 = o.nvfp4_quantize(d["x"], None)
    assert np.array_equal(s1, d["scale_1lvl"]) and np.array_equal(q1, d["q_1lvl"])
    pts = float(d["pts"][0])
    q2, s2 = o.nvfp4_quantize(d["x"], pts)
    assert np.array_equal(s2, d["scale_2lvl"]) and np.array_equal(q2, d["q_2lvl"])
    assert np.array_equal(o.nvfp4_dequant(q2, s2, pts), d["dq_2lvl_f32"])
    assert np.array_equal(o.f32_to_bf16(o.nvfp4_dequant(q2, s2, pts)), d["dq_2lvl"])
    assert np.array_equal(o.f32_to_bf16(o.nvfp4_dequant(q1, s1, None)), d["dq_1lvl"])
    assert np.array_equal(o.to_blocked(d["blk_in"]).reshape(-1), d["blk_out"].reshape(-1))


def test_int4_hqq_qparams_match_reference_within_solver_tolerance(golden_dir):
    """HQQ is a floating-point iterative solver: the oracle restates the reference's fp32 loop, but libm
    powf and the summation order inside torch.mean differ from torch's in the last bit; a last-bit difference in a
    group's zero point can move a code that sits on a rounding boundary, and a few groups then settle on the equivalent
    representation (q + 1, zero - scale).  Bars: scales bit-exact (they only depend on min/max); codes never more than
    one step apart and >= 99 % identical; zero points >= 97 % bit-identical, >= 99 % within 0.01 step, all within one
    step; mean reconstruction error |W - dequant| within 0.1 % of the reference's."""
    gold = load(golden_dir, "int4_hqq.npz")
    for name in ("g32", "g128", "g64_outlier"):
        w, g = gold[f"{name}__w"], int(gold[f"{name}__g"])
        q, s, z, iters = o.int4_hqq(w, g)
        qr, sr, zr = gold[f"{name}__q"], gold[f"{name}__s"], gold[f"{name}__z"]
        assert 1 <= iters <= 20
        assert np.array_equal(s, sr), name
        diff = np.abs(q.astype(np.int32) - qr.astype(np.int32))
        assert diff.max() <= 1 and (diff != 0).mean() <= 1e-2, (name, diff.max(), (diff != 0).mean())
        steps = np.abs(o.bf16_to_f32(z) - o.bf16_to_f32(zr)) / o.bf16_to_f32(sr)
        assert (z != zr).mean() <= 0.03 and (steps > 0.01).mean() <= 0.01 and steps.max() <= 1.0, (name, steps.max())

        def recon_err(qq, ss, zz):
            wf = o.bf16_to_f32(w).reshape(-1, g)
            deq = (qq.reshape(-1, g).astype(np.float32) - 8.0) * o.bf16_to_f32(ss).reshape(-1, 1) + o.bf16_to_f32(zz).reshape(-1, 1)
            return float(np.abs(wf - deq).mean())

        e, er = recon_err(q, s, z), recon_err(qr, sr, zr)
        assert abs(e - er) <= 1e-3 * er, (name, e, er)
