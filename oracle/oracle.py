"""numpy/ctypes front-end of the CPU oracle (oracle/ao_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (ao_b200/) never imports this.
bf16 arrays are numpy uint16 bit patterns; helpers convert from/to torch tensors.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    out = _DIR / "_build" / "libao_oracle.so"
    src = _DIR / "ao_oracle.c"
    if force or not out.exists() or out.stat().st_mtime < src.stat().st_mtime:
        r = subprocess.run(["make", "-C", str(_DIR)] + (["-B"] if force else []),
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stdout)
    return out


def lib():
    global _LIB
    if _LIB is None:
        path = _DIR / "_build" / "libao_oracle.so"
        if not path.exists():
            build()
        _LIB = C.CDLL(str(path))
        _LIB.ao_oracle_e4m3_to_f32.restype = C.c_float
        _LIB.ao_oracle_e4m3_to_f32.argtypes = [C.c_uint8]
        _LIB.ao_oracle_f32_to_e4m3.restype = C.c_uint8
        _LIB.ao_oracle_f32_to_e4m3.argtypes = [C.c_float]
        _LIB.ao_oracle_f32_to_e8m0_rceil.restype = C.c_uint8
        _LIB.ao_oracle_f32_to_e8m0_rceil.argtypes = [C.c_float]
        _LIB.ao_oracle_f32_to_e2m1.restype = C.c_uint8
        _LIB.ao_oracle_f32_to_e2m1.argtypes = [C.c_float]
        _LIB.ao_oracle_e2m1_to_f32.restype = C.c_float
        _LIB.ao_oracle_e2m1_to_f32.argtypes = [C.c_uint8]
    return _LIB


def _p(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


# ---- torch <-> numpy(bf16 bits) ------------------------------------------------------------
def bf16_bits(t) -> np.ndarray:
    """torch bf16 tensor (any device) -> numpy uint16 bit patterns."""
    import torch

    return t.detach().to("cpu").contiguous().view(torch.int16).numpy().view(np.uint16).copy()


def bf16_tensor(bits: np.ndarray):
    import torch

    return torch.from_numpy(bits.view(np.int16).copy()).view(torch.bfloat16)


def bf16_to_f32(bits: np.ndarray) -> np.ndarray:
    return (bits.astype(np.uint32) << 16).view(np.float32)


def f32_to_bf16(x: np.ndarray) -> np.ndarray:
    x = _c(x, np.float32)
    out = np.empty(x.shape, np.uint16)
    lib().ao_oracle_f32_to_bf16(_p(x), C.c_size_t(x.size), _p(out))
    return out


# ---- int4 tinygemm --------------------------------------------------------------------------
def int4_choose_qparams(w_bits: np.ndarray, g: int):
    N, K = w_bits.shape
    s = np.empty((N, K // g), np.uint16)
    z = np.empty((N, K // g), np.uint16)
    lib().ao_oracle_int4_choose_qparams(_p(_c(w_bits, np.uint16)), N, K, g, _p(s), _p(z))
    return s, z


def int4_quantize(w_bits, g, s, z):
    N, K = w_bits.shape
    q = np.empty((N, K), np.uint8)
    lib().ao_oracle_int4_quantize(_p(_c(w_bits, np.uint16)), N, K, g, _p(_c(s, np.uint16)), _p(_c(z, np.uint16)), _p(q))
    return q


def int4_hqq(w_bits: np.ndarray, g: int):
    """HQQ int4 codes [N, K] (one per byte) + tinygemm-convention scale / zero (bf16 bits, [N, K/g]); also the
    number of solver iterations run."""
    N, K = w_bits.shape
    q = np.empty((N, K), np.uint8)
    s = np.empty((N, K // g), np.uint16)
    z = np.empty((N, K // g), np.uint16)
    lib().ao_oracle_int4_hqq.restype = C.c_int
    iters = lib().ao_oracle_int4_hqq(_p(_c(w_bits, np.uint16)), N, K, g, _p(q), _p(s), _p(z))
    return q, s, z, int(iters)


def pack_scales_and_zeros(s, z):
    N, KG = s.shape
    out = np.empty((KG, N, 2), np.uint16)
    lib().ao_oracle_pack_scales_and_zeros(_p(_c(s, np.uint16)), _p(_c(z, np.uint16)), N, KG, _p(out))
    return out


def int4_pack_tile4d(q: np.ndarray, ikt: int = 8):
    N, K = q.shape
    out = np.empty((N // 8, K // (ikt * 16), 32, ikt // 2), np.int32)
    lib().ao_oracle_int4_pack_tile4d(_p(_c(q, np.uint8)), N, K, ikt, _p(out))
    return out


def int4_unpack_tile4d(qd: np.ndarray):
    n8, kt, _, wpl = qd.shape
    ikt = wpl * 2
    N, K = n8 * 8, kt * ikt * 16
    q = np.empty((N, K), np.uint8)
    lib().ao_oracle_int4_unpack_tile4d(_p(_c(qd, np.int32)), N, K, ikt, _p(q))
    return q


def int4_prepack_bytes(q: np.ndarray) -> np.ndarray:
    """byte = q[n,2j]<<4 | q[n,2j+1] (int4_tile_packed_to_4d_tensor.py:199)."""
    return ((q[:, ::2].astype(np.uint8) << 4) | q[:, 1::2].astype(np.uint8)).astype(np.uint8)


def int4_dequant(q: np.ndarray, sz: np.ndarray, g: int) -> np.ndarray:
    N, K = q.shape
    out = np.empty((N, K), np.uint16)
    lib().ao_oracle_int4_dequant(_p(_c(q, np.uint8)), _p(_c(sz, np.uint16)), N, K, g, _p(out))
    return out


def linear_f32(x: np.ndarray, w: np.ndarray, bias=None) -> np.ndarray:
    """y = x @ w.T + bias with exact products and double accumulation -> float32."""
    x = _c(x, np.float32)
    w = _c(w, np.float32)
    M, K = x.shape
    N = w.shape[0]
    y = np.empty((M, N), np.float32)
    b = _c(bias, np.float32) if bias is not None else None
    lib().ao_oracle_linear_f32(_p(x), M, K, _p(w), N, _p(b) if b is not None else None, _p(y))
    return y


def int4_linear(x_bits, qd, sz, g, bias_bits=None) -> np.ndarray:
    M, K = x_bits.shape
    N = qd.shape[0] * 8
    y = np.empty((M, N), np.uint16)
    b = _c(bias_bits, np.uint16) if bias_bits is not None else None
    lib().ao_oracle_int4_linear(_p(_c(x_bits, np.uint16)), M, K, _p(_c(qd, np.int32)), _p(_c(sz, np.uint16)), g, N,
                                _p(b) if b is not None else None, _p(y))
    return y


# ---- int8 -----------------------------------------------------------------------------------
def int8_quantize_rowwise(x_bits):
    M, K = x_bits.shape
    q = np.empty((M, K), np.int8)
    s = np.empty((M,), np.float32)
    lib().ao_oracle_int8_quantize_rowwise(_p(_c(x_bits, np.uint16)), M, K, _p(q), _p(s))
    return q, s


def int8_mm(a, b):
    M, K = a.shape
    N = b.shape[0]
    c = np.empty((M, N), np.int32)
    lib().ao_oracle_int8_mm(_p(_c(a, np.int8)), M, K, _p(_c(b, np.int8)), N, _p(c))
    return c


def int8_epilogue(acc, sx, sw, bias_bits=None):
    M, N = acc.shape
    y = np.empty((M, N), np.uint16)
    b = _c(bias_bits, np.uint16) if bias_bits is not None else None
    lib().ao_oracle_int8_epilogue(_p(_c(acc, np.int32)), M, N, _p(_c(sx, np.float32)), _p(_c(sw, np.float32)),
                                  _p(b) if b is not None else None, _p(y))
    return y


# ---- fp8 ------------------------------------------------------------------------------------
def fp8_quantize_rowwise(x_bits):
    M, K = x_bits.shape
    q = np.empty((M, K), np.uint8)
    s = np.empty((M,), np.float32)
    lib().ao_oracle_fp8_quantize_rowwise(_p(_c(x_bits, np.uint16)), M, K, _p(q), _p(s))
    return q, s


def e4m3_to_f32(q: np.ndarray) -> np.ndarray:
    q = _c(q, np.uint8)
    out = np.empty(q.shape, np.float32)
    lib().ao_oracle_e4m3_array_to_f32(_p(q), C.c_size_t(q.size), _p(out))
    return out


# ---- mxfp8 ----------------------------------------------------------------------------------
def mxfp8_quantize(x_bits):
    M, K = x_bits.shape
    q = np.empty((M, K), np.uint8)
    s = np.empty((M, K // 32), np.uint8)
    lib().ao_oracle_mxfp8_quantize(_p(_c(x_bits, np.uint16)), M, K, _p(q), _p(s))
    return q, s


def to_blocked(a: np.ndarray) -> np.ndarray:
    a = _c(a, np.uint8)
    H, W = a.shape
    out = np.empty((32 * ((H + 127) // 128), 16 * ((W + 3) // 4)), np.uint8)
    lib().ao_oracle_to_blocked(_p(a), H, W, _p(out))
    return out


def from_blocked(a: np.ndarray, H: int, W: int) -> np.ndarray:
    out = np.empty((H, W), np.uint8)
    lib().ao_oracle_from_blocked(_p(_c(a, np.uint8)), H, W, _p(out))
    return out


def mxfp8_dequant(q, s) -> np.ndarray:
    M, K = q.shape
    out = np.empty((M, K), np.float32)
    lib().ao_oracle_mxfp8_dequant(_p(_c(q, np.uint8)), _p(_c(s, np.uint8)), M, K, _p(out))
    return out


# ---- nvfp4 ----------------------------------------------------------------------------------
def nvfp4_quantize(x_bits, pts=None):
    M, K = x_bits.shape
    q = np.empty((M, K // 2), np.uint8)
    s = np.empty((M, K // 16), np.uint8)
    p = C.byref(C.c_float(float(pts))) if pts is not None else None
    lib().ao_oracle_nvfp4_quantize(_p(_c(x_bits, np.uint16)), M, K, p, _p(q), _p(s))
    return q, s


def nvfp4_dequant(q, s, pts=None) -> np.ndarray:
    M, K2 = q.shape
    K = K2 * 2
    out = np.empty((M, K), np.float32)
    p = C.byref(C.c_float(float(pts))) if pts is not None else None
    lib().ao_oracle_nvfp4_dequant(_p(_c(q, np.uint8)), _p(_c(s, np.uint8)), p, M, K, _p(out))
    return out


def sqnr_db(ref: np.ndarray, out: np.ndarray) -> float:
    """20*log10(|ref| / |ref-out|) (quantization/utils.py:59-62 compute_error)."""
    ref = ref.astype(np.float64)
    out = out.astype(np.float64)
    den = np.linalg.norm(ref - out)
    if den == 0:
        return float("inf")
    return float(20 * np.log10(np.linalg.norm(ref) / den))
