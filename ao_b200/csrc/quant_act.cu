// Dynamic activation quantisation prologues (bit-exact restatements of the reference's torch
// ops, one fused kernel each instead of the 2-6 eager kernels the reference launches):
//   int8 per-token symmetric : Int8Tensor.from_hp(x, PerRow())  int8_tensor.py:176-248,
//                              quant_primitives.py:1487-1583 / :424-485
//   e4m3 per-token           : _choose_scale_float8 + _quantize_affine_float8
//                              quant_primitives.py:2172-2287 (float8_tensor.py:235-242)
//   mxfp8 RCEIL block-32     : to_mx  mx_formats/mx_tensor.py:228-409, :111-225
//   nvfp4 block-16           : nvfp4_quantize  mx_formats/nvfp4_tensor.py:772-854
// and the 128x4 -> 32x16 scale swizzle (mx_formats/utils.py:31-70) fused into the writers.
// Inputs are bf16 [M,K]; HBM-bound elementwise/reduction work: 16-byte vector loads, one
// pass for the reduction and one for the cast (the row stays in L1/L2).
#include <cuda_bf16.h>
#include <cuda_fp4.h>
#include <cuda_fp8.h>

#include "common.h"
#include "ptx.cuh"

namespace ao {

__device__ __forceinline__ float bf16_round(float v) {
  return __bfloat162float(__float2bfloat16_rn(v));
}
__device__ __forceinline__ float block_reduce_max(float v, float* sh) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  const int nw = blockDim.x >> 5;
  v = (l < nw) ? sh[l] : 0.f;
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  return v;
}
// NaN-propagating abs-max like torch.amax(abs(x))
__device__ __forceinline__ float nanmax(float a, float b) {
  return (a != a || b != b) ? __int_as_float(0x7fc00000) : fmaxf(a, b);
}

// ---------------------------------------------------------------- int8 / fp8 rowwise
template <int MODE>  // 0 = int8, 1 = e4m3
__global__ void __launch_bounds__(256) quant_rowwise_kernel(const __nv_bfloat16* __restrict__ x, int ldx,
                                                            int K, uint8_t* __restrict__ q,
                                                            float* __restrict__ scale) {
  __shared__ float sh[8];
  // PDL: let the linear that consumes this output become resident and prefetch its weights now; our own input may
  // be the previous kernel's output, so wait for it before the first read
  pdl_launch_dependents();
  pdl_wait();
  const int m = blockIdx.x;
  const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)m * ldx);   // row pitch ldx >= K (a column slice)
  const int nv = K / 8;
  float amax = 0.f;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    const uint4 v = xr[i];
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(h[j]);
      amax = fmaxf(amax, fmaxf(fabsf(f.x), fabsf(f.y)));
    }
  }
  amax = block_reduce_max(amax, sh);
  float s;
  if (MODE == 0) {
    s = bf16_round(amax / 127.5f);                 // division happens in the input dtype (bf16)
    s = fmaxf(s, 1.1920928955078125e-07f);         // eps = finfo(float32).eps
  } else {
    s = bf16_round(amax / 448.0f);                 // no eps (reference has none)
  }
  if (threadIdx.x == 0) scale[m] = s;
  const float inv = 1.0f / s;
  uint2* qr = reinterpret_cast<uint2*>(q + (size_t)m * K);
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    const uint4 v = xr[i];
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
    uint8_t o[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(h[j]);
      if (MODE == 0) {
        const float a = fminf(fmaxf(rintf(f.x * inv), -128.f), 127.f);
        const float b = fminf(fmaxf(rintf(f.y * inv), -128.f), 127.f);
        o[2 * j] = (uint8_t)(int8_t)(int)a;
        o[2 * j + 1] = (uint8_t)(int8_t)(int)b;
      } else {
        float a = f.x / s, b = f.y / s;
        a = fminf(fmaxf(a, -448.f), 448.f);  // fminf/fmaxf drop NaN like torch.clamp? no: keep NaN
        b = fminf(fmaxf(b, -448.f), 448.f);
        if (s == 0.f) { a = __int_as_float(0x7fc00000); b = a; }  // 0/0 = NaN in the reference
        o[2 * j] = (uint8_t)__nv_cvt_float_to_fp8(a, __NV_SATFINITE, __NV_E4M3);
        o[2 * j + 1] = (uint8_t)__nv_cvt_float_to_fp8(b, __NV_SATFINITE, __NV_E4M3);
      }
    }
    qr[i] = *reinterpret_cast<const uint2*>(o);
  }
}

// Same arithmetic, the row held in REGISTERS between the abs-max pass and the cast pass (K <= 16384): `tpr` threads
// per row (32 .. 256, whole warps), 256 / tpr rows per CTA, up to 8 x 16-byte loads per thread all in flight before the
// first use, packed hardware converts (cvt.rn.satfinite.e4m3x2.f32 / cvt.rni.sat.s8.f32).  The 2-pass kernel above
// stays for longer rows.
__device__ __forceinline__ uint32_t pack_s8x4(float a, float b, float c, float d) {
  int ia, ib, ic, id;
  asm("cvt.rni.sat.s8.f32 %0, %1;" : "=r"(ia) : "f"(a));   // round-to-nearest-even + clamp to [-128, 127]
  asm("cvt.rni.sat.s8.f32 %0, %1;" : "=r"(ib) : "f"(b));
  asm("cvt.rni.sat.s8.f32 %0, %1;" : "=r"(ic) : "f"(c));
  asm("cvt.rni.sat.s8.f32 %0, %1;" : "=r"(id) : "f"(d));
  return (uint32_t)(ia & 0xff) | ((uint32_t)(ib & 0xff) << 8) | ((uint32_t)(ic & 0xff) << 16) | ((uint32_t)id << 24);
}
__device__ __forceinline__ uint32_t pack_e4m3x4(float a, float b, float c, float d) {
  const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
  const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E4M3);
  return lo | (hi << 16);
}

template <int MODE>  // 0 = int8, 1 = e4m3
__global__ void __launch_bounds__(256, 3) quant_rowwise_reg_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int M, int K,
                                                                int tpr, uint8_t* __restrict__ q,
                                                                float* __restrict__ scale) {
  constexpr int VPT = 8;
  __shared__ float sh[8];
  pdl_launch_dependents();
  pdl_wait();
  const int row_in_cta = threadIdx.x / tpr, t = threadIdx.x % tpr;
  const int m = blockIdx.x * (256 / tpr) + row_in_cta;
  const bool active = m < M;
  const int nv = K / 8;
  const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)(active ? m : 0) * ldx);
  uint4 v[VPT];
#pragma unroll
  for (int j = 0; j < VPT; ++j) {
    const int i = t + j * tpr;
    v[j] = (active && i < nv) ? xr[i] : make_uint4(0u, 0u, 0u, 0u);
  }
  float amax = 0.f;
#pragma unroll
  for (int j = 0; j < VPT; ++j) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v[j]);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __bfloat1622float2(h[e]);
      amax = fmaxf(amax, fmaxf(fabsf(f.x), fabsf(f.y)));
    }
  }
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if (tpr > 32) {   // (uniform) the row spans tpr / 32 warps
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) sh[w] = amax;
    __syncthreads();
    const int w0 = row_in_cta * (tpr >> 5);
    amax = sh[w0];
    for (int i = 1; i < (tpr >> 5); ++i) amax = fmaxf(amax, sh[w0 + i]);
  }
  float s;
  if (MODE == 0) s = fmaxf(bf16_round(amax / 127.5f), 1.1920928955078125e-07f);
  else s = bf16_round(amax / 448.0f);
  if (active && t == 0) scale[m] = s;
  const float inv = 1.0f / s;
  uint2* qr = reinterpret_cast<uint2*>(q + (size_t)(active ? m : 0) * K);
#pragma unroll
  for (int j = 0; j < VPT; ++j) {
    const int i = t + j * tpr;
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v[j]);
    float f[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 t2 = __bfloat1622float2(h[e]);
      f[2 * e] = t2.x;
      f[2 * e + 1] = t2.y;
    }
    uint2 o;
    if (MODE == 0) {
      o.x = pack_s8x4(f[0] * inv, f[1] * inv, f[2] * inv, f[3] * inv);
      o.y = pack_s8x4(f[4] * inv, f[5] * inv, f[6] * inv, f[7] * inv);
    } else {
      // x / s with the reference's IEEE rounding.  The row's scale is uniform, so its reciprocal is computed once and
      // every quotient costs a multiply and ONE residual correction (q = x*r; q += (x - q*s) * r: what div.rn.f32
      // itself does after refining the reciprocal; exact residual through the FMA).  Valid while nothing can leave the
      // normal range: |x| <= amax ~ 448 s, so it is enough that s is far from 0 / inf; otherwise the plain division.
      if (s >= 0x1p-64f && s <= 0x1p64f) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float q0 = f[e] * inv;
          f[e] = fminf(fmaxf(fmaf(fmaf(-q0, s, f[e]), inv, q0), -448.f), 448.f);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          f[e] = fminf(fmaxf(f[e] / s, -448.f), 448.f);
          if (s == 0.f) f[e] = __int_as_float(0x7fc00000);   // 0/0 = NaN in the reference
        }
      }
      o.x = pack_e4m3x4(f[0], f[1], f[2], f[3]);
      o.y = pack_e4m3x4(f[4], f[5], f[6], f[7]);
    }
    if (active && i < nv) qr[i] = o;
  }
}

// ---------------------------------------------------------------- producer-fused rowwise quantizers
// the activation quantization of a dynamic-activation linear fused with the op that produces the
// activations, so the bf16 activations never travel to HBM and back:
//   PRO 1  RMSNorm   y = bf16(w * bf16(x_f32 * rsqrt(mean(x_f32^2) + eps)))      (HF LlamaRMSNorm: fp32 statistics,
//                                                                                  cast to the input dtype, * weight)
//   PRO 2  SiLU-mul  y = bf16(bf16(silu_f32(g)) * u)                              (HF LlamaMLP: act_fn(gate) * up)
// followed by exactly quant_rowwise_kernel's arithmetic on y (MODE 0 int8 per token, MODE 1 e4m3 per token).  One
// CTA per token; the row of y is kept in shared memory between the abs-max pass and the cast pass.  (A register-resident
// variant like quant_rowwise_reg_kernel was measured SLOWER here -- 128 registers, two CTAs per SM: 2.0 / 1.4 TB/s against
// 2.5 / 2.6 TB/s, profiles/r02_call_p.log -- and removed.)
__device__ __forceinline__ float block_reduce_sum(float v, float* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  const int nw = blockDim.x >> 5;
  v = (l < nw) ? sh[l] : 0.f;
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  return v;
}

template <int MODE, int PRO>
__global__ void __launch_bounds__(256) fused_rowwise_kernel(const __nv_bfloat16* __restrict__ a, int lda,
                                                            const __nv_bfloat16* __restrict__ b, int ldb, float eps,
                                                            int K, uint8_t* __restrict__ q, float* __restrict__ scale) {
  extern __shared__ uint4 yrow[];   // K / 8 vectors of 8 bf16
  __shared__ float sh[8];
  pdl_launch_dependents();
  pdl_wait();
  const int m = blockIdx.x;
  const uint4* ar = reinterpret_cast<const uint4*>(a + (size_t)m * lda);
  const uint4* br = reinterpret_cast<const uint4*>(PRO == 1 ? b : b + (size_t)m * ldb);   // weight[K] or up[m, :]
  const int nv = K / 8;
  float rstd = 0.f;
  if (PRO == 1) {
    float ss = 0.f;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
      const uint4 v = ar[i];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(h[j]);
        ss += f.x * f.x + f.y * f.y;
      }
    }
    ss = block_reduce_sum(ss, sh);
    rstd = rsqrtf(ss / (float)K + eps);
  }
  float amax = 0.f;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    const uint4 va = ar[i], vb = br[i];
    const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&va);
    const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&vb);
    uint4 out;
    __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&out);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fa = __bfloat1622float2(ha[j]), fb = __bfloat1622float2(hb[j]);
      float y0, y1;
      if (PRO == 1) {
        y0 = bf16_round(fb.x * bf16_round(fa.x * rstd));
        y1 = bf16_round(fb.y * bf16_round(fa.y * rstd));
      } else {
        y0 = bf16_round(bf16_round(fa.x / (1.f + expf(-fa.x))) * fb.x);
        y1 = bf16_round(bf16_round(fa.y / (1.f + expf(-fa.y))) * fb.y);
      }
      ho[j] = __floats2bfloat162_rn(y0, y1);
      amax = fmaxf(amax, fmaxf(fabsf(y0), fabsf(y1)));
    }
    yrow[i] = out;
  }
  amax = block_reduce_max(amax, sh);   // (its __syncthreads also publish yrow)
  float s;
  if (MODE == 0) {
    s = fmaxf(bf16_round(amax / 127.5f), 1.1920928955078125e-07f);
  } else {
    s = bf16_round(amax / 448.0f);
  }
  if (threadIdx.x == 0) scale[m] = s;
  const float inv = 1.0f / s;
  uint2* qr = reinterpret_cast<uint2*>(q + (size_t)m * K);
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    const uint4 v = yrow[i];
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
    uint8_t o[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(h[j]);
      if (MODE == 0) {
        o[2 * j] = (uint8_t)(int8_t)(int)fminf(fmaxf(rintf(f.x * inv), -128.f), 127.f);
        o[2 * j + 1] = (uint8_t)(int8_t)(int)fminf(fmaxf(rintf(f.y * inv), -128.f), 127.f);
      } else {
        float x0 = fminf(fmaxf(f.x / s, -448.f), 448.f), x1 = fminf(fmaxf(f.y / s, -448.f), 448.f);
        if (s == 0.f) { x0 = __int_as_float(0x7fc00000); x1 = x0; }  // 0/0 = NaN in the reference
        o[2 * j] = (uint8_t)__nv_cvt_float_to_fp8(x0, __NV_SATFINITE, __NV_E4M3);
        o[2 * j + 1] = (uint8_t)__nv_cvt_float_to_fp8(x1, __NV_SATFINITE, __NV_E4M3);
      }
    }
    qr[i] = *reinterpret_cast<const uint2*>(o);
  }
}

// ---------------------------------------------------------------- block-scaled formats
__device__ __forceinline__ size_t blocked_index(int r, int c, int col_blocks) {
  // mx_formats/utils.py:31-70: tile (r/128, c/4) of 512 bytes; (r%32)*16 + ((r%128)/32)*4 + c%4
  return ((size_t)(r >> 7) * col_blocks + (c >> 2)) * 512 + (r & 31) * 16 + ((r & 127) >> 5) * 4 + (c & 3);
}

__device__ __forceinline__ uint8_t e8m0_rceil(float v) {
  const uint32_t u = __float_as_uint(v);
  if (!isfinite(v)) return 0xff;
  const uint32_t be = (u >> 23) & 0xff, man = u & 0x7fffff;
  const uint32_t up = (be == 0) ? (man > 0x400000u) : (man != 0);
  return (uint8_t)(be + up);
}
__device__ __forceinline__ float e8m0_recip(uint8_t e) {
  const uint8_t r = (uint8_t)(254 - (int)e);
  uint32_t bits = (uint32_t)r << 23;
  if (r == 0) bits = 0x00400000u;
  if (r == 0xff) bits = 0x7f800001u;
  return __uint_as_float(bits);
}

// Zero entries of the PADDED blocked scale grid (rows to a multiple of 128, blocks to a multiple of 4): written by the
// quantizer itself so no separate memset launch is needed (the GEMM multiplies them with TMA's zero fill: they must not
// be NaN).  Called by every thread of the grid; the padding is (Mp - M) x nbp + M x (nbp - nb) entries.
__device__ __forceinline__ void zero_scale_padding(uint8_t* sc, int M, int nb, size_t tid, size_t nthreads) {
  const int nbp = (nb + 3) & ~3, Mp = (M + 127) & ~127;
  const size_t pad_rows = (size_t)(Mp - M) * nbp, pad_cols = (size_t)M * (nbp - nb);
  for (size_t i = tid; i < pad_rows + pad_cols; i += nthreads) {
    int m, kb;
    if (i < pad_rows) { m = M + (int)(i / nbp); kb = (int)(i % nbp); }
    else { const size_t j = i - pad_rows; m = (int)(j / (nbp - nb)); kb = nb + (int)(j % (nbp - nb)); }
    sc[blocked_index(m, kb, nbp / 4)] = 0;
  }
}

// mxfp8: one thread per 32-element block (64 contiguous bytes in, 32 out; a warp covers 2 KB of a row).  Measured
// against a four-lanes-per-block variant with fully coalesced 16-byte accesses (profiles/r02_call_l.log: 3.9 TB/s):
// the per-block work (abs-max, scale, reciprocal) is done once instead of four times and needs no shuffles, which
// matters more than the access pattern -- the kernel is issue-bound, not sector-bound.  The NaN-propagating abs-max
// runs on the bf16 BIT PATTERNS: |x| as an unsigned integer orders like the value and every NaN sorts above inf.
constexpr int BQ_THREADS = 128;
__global__ void __launch_bounds__(BQ_THREADS) mxfp8_quant_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int M, int K,
                                                                 uint8_t* __restrict__ q, uint8_t* __restrict__ sc,
                                                                 int swizzled) {
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t nb = K / 32;
  const uint32_t idx = blockIdx.x * BQ_THREADS + threadIdx.x;   // (the launcher checks M * nb < 2^32)
  if (idx < (uint32_t)M * nb) {
    const uint32_t m = idx / nb, kb = idx - m * nb;
    const uint4* src = reinterpret_cast<const uint4*>(x + (size_t)m * ldx + kb * 32);
    uint4 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = src[i];
    uint32_t mx = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      mx = __vmaxu2(mx, v[i].x & 0x7FFF7FFFu);
      mx = __vmaxu2(mx, v[i].y & 0x7FFF7FFFu);
      mx = __vmaxu2(mx, v[i].z & 0x7FFF7FFFu);
      mx = __vmaxu2(mx, v[i].w & 0x7FFF7FFFu);
    }
    const uint32_t abits = max(mx & 0xFFFFu, mx >> 16);
    const float amax = __uint_as_float(abits << 16);          // NaN when any element was NaN
    const uint8_t e8 = e8m0_rceil(amax * (float)(1.0 / 448.0));
    const float r = e8m0_recip(e8);
    if (swizzled) sc[blocked_index(m, kb, (nb + 3) / 4)] = e8;
    else sc[(size_t)m * nb + kb] = e8;
    uint4 o[2];
    uint32_t* ow = reinterpret_cast<uint32_t*>(o);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v[i]);
      const float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
      const float2 f2 = __bfloat1622float2(h[2]), f3 = __bfloat1622float2(h[3]);
      ow[2 * i] = pack_e4m3x4(f0.x * r, f0.y * r, f1.x * r, f1.y * r);
      ow[2 * i + 1] = pack_e4m3x4(f2.x * r, f2.y * r, f3.x * r, f3.y * r);
    }
    uint4* dst = reinterpret_cast<uint4*>(q + (size_t)m * K + kb * 32);
    dst[0] = o[0];
    dst[1] = o[1];
  }
  if (swizzled) zero_scale_padding(sc, M, nb, (size_t)blockIdx.x * BQ_THREADS + threadIdx.x, (size_t)gridDim.x * BQ_THREADS);
}

// e2m1 RNE, saturating (custom_fp_utils.py:27-146): thresholds are the midpoints, ties to even
__device__ __forceinline__ uint32_t f32_to_e2m1(float f) {
  const uint32_t s = (__float_as_uint(f) >> 31) << 3;
  const float a = fabsf(f);
  uint32_t c;
  if (!(a < 5.0f)) c = (a == 5.0f) ? 6 : 7;        // 5.0 ties to 4 (code 6, even); NaN -> 7
  else if (a >= 3.5f) c = 6;                        // 3.5 ties to 4
  else if (a > 2.5f) c = 5;                         // 2.5 ties to 2 (code 4)
  else if (a >= 1.75f) c = 4;                       // 1.75 ties to 2
  else if (a > 1.25f) c = 3;                        // 1.25 ties to 1
  else if (a >= 0.75f) c = 2;                       // 0.75 ties to 1
  else if (a > 0.25f) c = 1;                        // 0.25 ties to 0
  else c = 0;
  return s | c;
}

// nvfp4: one thread per 16-element block (32 contiguous bytes in, 8 out); e2m1 pairs by cvt.rn.satfinite.e2m1x2.f32
// (RNE, saturating: the reference's rounding, custom_fp_utils.py:27-146; round 1 used a seven-way comparison chain per
// element), NaN through that chain (the hardware convert canonicalises the sign).
__device__ __forceinline__ uint32_t e2m1_pair(float a, float b) {
  a = fminf(fmaxf(a, -6.f), 6.f);
  b = fminf(fmaxf(b, -6.f), 6.f);
  if (a != a || b != b) return f32_to_e2m1(a) | (f32_to_e2m1(b) << 4);
  return (uint32_t)__nv_cvt_float2_to_fp4x2(make_float2(a, b), __NV_E2M1, cudaRoundNearest) & 0xffu;   // a in the LOW nibble
}
__global__ void __launch_bounds__(BQ_THREADS) nvfp4_quant_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int M, int K,
                                                                 const float* __restrict__ pts, uint8_t* __restrict__ q,
                                                                 uint8_t* __restrict__ sc, int swizzled) {
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t nb = K / 16;
  const uint32_t idx = blockIdx.x * BQ_THREADS + threadIdx.x;   // (the launcher checks M * nb < 2^32)
  if (idx < (uint32_t)M * nb) {
    const uint32_t m = idx / nb, kb = idx - m * nb;
    const uint4* src = reinterpret_cast<const uint4*>(x + (size_t)m * ldx + kb * 16);
    const uint4 v0 = src[0], v1 = src[1];
    float f[16];
    float amax = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(i == 0 ? &v0 : &v1);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = __bfloat1622float2(h[j]);
        f[i * 8 + 2 * j] = t.x;
        f[i * 8 + 2 * j + 1] = t.y;
        amax = fmaxf(amax, fmaxf(fabsf(t.x), fabsf(t.y)));
      }
    }
    const float bs = amax / 6.0f;
    float recip;
    uint8_t b8;
    if (pts == nullptr) {
      const float c = fminf(fmaxf(bs, 0.015625f), 448.f);
      b8 = (uint8_t)__nv_cvt_float_to_fp8(c, __NV_SATFINITE, __NV_E4M3);
      const float bf = __half2float(__half(__nv_cvt_fp8_to_halfraw(b8, __NV_E4M3)));
      recip = 1.0f / bf;
    } else {
      const float p = *pts;
      const float c = fminf(fmaxf(bs / p, 0.015625f), 448.f);
      b8 = (uint8_t)__nv_cvt_float_to_fp8(c, __NV_SATFINITE, __NV_E4M3);
      const float bf = __half2float(__half(__nv_cvt_fp8_to_halfraw(b8, __NV_E4M3)));
      recip = (1.0f / p) / bf;
    }
    if (swizzled) sc[blocked_index(m, kb, (nb + 3) / 4)] = b8;
    else sc[(size_t)m * nb + kb] = b8;
    uint2 o;
    o.x = o.y = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {   // even k in the LOW nibble
      o.x |= e2m1_pair(f[2 * e] * recip, f[2 * e + 1] * recip) << (8 * e);
      o.y |= e2m1_pair(f[8 + 2 * e] * recip, f[8 + 2 * e + 1] * recip) << (8 * e);
    }
    *reinterpret_cast<uint2*>(q + (size_t)m * (K / 2) + kb * 8) = o;
  }
  if (swizzled) zero_scale_padding(sc, M, nb, (size_t)blockIdx.x * BQ_THREADS + threadIdx.x, (size_t)gridDim.x * BQ_THREADS);
}

}  // namespace ao

using namespace ao;

template <int MODE, int PRO>
static int launch_fused(const uint16_t* a, int lda, const uint16_t* b, int ldb, float eps, int M, int K, uint8_t* q, float* scale, void* stream) {
  auto kern = fused_rowwise_kernel<MODE, PRO>;
  const size_t smem = (size_t)K * 2;
  AO_CUDA_CHECK(ensure_dynamic_smem(reinterpret_cast<const void*>(kern), 96 * 1024));
  AO_CUDA_CHECK(ao::launch(kern, dim3(M), dim3(256), smem, reinterpret_cast<cudaStream_t>(stream), pdl_enabled(),
                           reinterpret_cast<const __nv_bfloat16*>(a), lda, reinterpret_cast<const __nv_bfloat16*>(b), ldb, eps, K, q, scale));
  return AO_OK;
}

template <int MODE>
static int launch_rowwise(const uint16_t* x, int ldx, int M, int K, uint8_t* q, float* scale, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x);
  if (K <= 16384) {
    // threads per row: the fewest whole warps that hold the row in 8 vectors of 8 elements per thread
    int tpr = 32;
    while (tpr * 64 < K) tpr *= 2;
    AO_CUDA_CHECK(ao::launch(quant_rowwise_reg_kernel<MODE>, dim3((unsigned)ceil_div(M, 256 / tpr)), dim3(256), 0, st,
                             pdl_enabled(), xb, ldx, M, K, tpr, q, scale));
  } else {
    AO_CUDA_CHECK(ao::launch(quant_rowwise_kernel<MODE>, dim3(M), dim3(256), 0, st, pdl_enabled(), xb, ldx, K, q, scale));
  }
  return AO_OK;
}

static int check_ld(const char* what, const void* x, int ldx, int K) {
  AO_REQUIRE(ldx >= K && ldx % 8 == 0, "%s: ldx=%d must be >= K=%d and a multiple of 8", what, ldx, K);
  AO_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "%s: x must be 16-byte aligned", what);
  return AO_OK;
}

extern "C" int ao_int8_quantize_rowwise_ld(const uint16_t* x, int ldx, int M, int K, int8_t* q, float* scale, void* stream) {
  AO_REQUIRE(M >= 0 && K > 0 && K % 8 == 0, "int8 quantize: bad sizes M=%d K=%d (K%%8==0)", M, K);
  if (M == 0) return AO_OK;
  AO_REQUIRE(x && q && scale, "int8 quantize: null pointer");
  if (int rc = check_ld("int8 quantize", x, ldx, K)) return rc;
  return launch_rowwise<0>(x, ldx, M, K, reinterpret_cast<uint8_t*>(q), scale, stream);
}
extern "C" int ao_int8_quantize_rowwise(const uint16_t* x, int M, int K, int8_t* q, float* scale, void* stream) {
  return ao_int8_quantize_rowwise_ld(x, K, M, K, q, scale, stream);
}

extern "C" int ao_fp8_quantize_rowwise_ld(const uint16_t* x, int ldx, int M, int K, uint8_t* q, float* scale, void* stream) {
  AO_REQUIRE(M >= 0 && K > 0 && K % 8 == 0, "fp8 quantize: bad sizes M=%d K=%d (K%%8==0)", M, K);
  if (M == 0) return AO_OK;
  AO_REQUIRE(x && q && scale, "fp8 quantize: null pointer");
  if (int rc = check_ld("fp8 quantize", x, ldx, K)) return rc;
  return launch_rowwise<1>(x, ldx, M, K, q, scale, stream);
}
extern "C" int ao_fp8_quantize_rowwise(const uint16_t* x, int M, int K, uint8_t* q, float* scale, void* stream) {
  return ao_fp8_quantize_rowwise_ld(x, K, M, K, q, scale, stream);
}

extern "C" int ao_mxfp8_quantize_ld(const uint16_t* x, int ldx, int M, int K, uint8_t* q, uint8_t* scale_e8m0, int swizzled,
                                    void* stream) {
  AO_REQUIRE(M >= 0 && K > 0 && K % 32 == 0, "mxfp8 quantize: K=%d must be a multiple of 32 (mx_tensor.py:244-246)", K);
  if (M == 0) return AO_OK;
  AO_REQUIRE(x && q && scale_e8m0, "mxfp8 quantize: null pointer");
  if (int rc = check_ld("mxfp8 quantize", x, ldx, K)) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t total = (size_t)M * (K / 32);   // blocks = threads
  AO_REQUIRE(total < ((size_t)1 << 32) - BQ_THREADS, "mxfp8 quantize: M*K too large (%d x %d)", M, K);
  AO_CUDA_CHECK(ao::launch(mxfp8_quant_kernel, dim3((unsigned)((total + BQ_THREADS - 1) / BQ_THREADS)), dim3(BQ_THREADS), 0, st, pdl_enabled(),
                           reinterpret_cast<const __nv_bfloat16*>(x), ldx, M, K, q, scale_e8m0, swizzled));
  return AO_OK;
}
extern "C" int ao_mxfp8_quantize(const uint16_t* x, int M, int K, uint8_t* q, uint8_t* scale_e8m0, int swizzled, void* stream) {
  return ao_mxfp8_quantize_ld(x, K, M, K, q, scale_e8m0, swizzled, stream);
}

extern "C" int ao_nvfp4_quantize_ld(const uint16_t* x, int ldx, int M, int K, const float* per_tensor_scale, uint8_t* q,
                                    uint8_t* scale_e4m3, int swizzled, void* stream) {
  AO_REQUIRE(M >= 0 && K > 0 && K % 16 == 0, "nvfp4 quantize: K=%d must be a multiple of 16", K);
  if (M == 0) return AO_OK;
  AO_REQUIRE(x && q && scale_e4m3, "nvfp4 quantize: null pointer");
  if (int rc = check_ld("nvfp4 quantize", x, ldx, K)) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t total = (size_t)M * (K / 16);   // blocks = threads
  AO_REQUIRE(total < ((size_t)1 << 32) - BQ_THREADS, "nvfp4 quantize: M*K too large (%d x %d)", M, K);
  AO_CUDA_CHECK(ao::launch(nvfp4_quant_kernel, dim3((unsigned)((total + BQ_THREADS - 1) / BQ_THREADS)), dim3(BQ_THREADS), 0, st, pdl_enabled(),
                           reinterpret_cast<const __nv_bfloat16*>(x), ldx, M, K, per_tensor_scale, q, scale_e4m3, swizzled));
  return AO_OK;
}
extern "C" int ao_nvfp4_quantize(const uint16_t* x, int M, int K, const float* per_tensor_scale, uint8_t* q,
                                 uint8_t* scale_e4m3, int swizzled, void* stream) {
  return ao_nvfp4_quantize_ld(x, K, M, K, per_tensor_scale, q, scale_e4m3, swizzled, stream);
}

// RMSNorm -> per-token quantization.  x bf16 [M, K] with row pitch ldx, weight bf16 [K];
// fmt 0 = int8 (scale = max(bf16(amax/127.5), eps32)), 1 = e4m3 (scale = bf16(amax/448)); q [M, K] bytes, scale f32 [M].
extern "C" int ao_rmsnorm_quantize_rowwise(const uint16_t* x, int ldx, const uint16_t* weight, float eps, int M, int K, int fmt,
                                           uint8_t* q, float* scale, void* stream) {
  AO_REQUIRE(M >= 0 && K > 0 && K % 8 == 0 && K <= 48 * 1024, "rmsnorm quantize: bad sizes M=%d K=%d (K%%8==0, K<=49152)", M, K);
  AO_REQUIRE(fmt == 0 || fmt == 1, "rmsnorm quantize: fmt=%d (0 = int8, 1 = e4m3)", fmt);
  if (M == 0) return AO_OK;
  AO_REQUIRE(x && weight && q && scale, "rmsnorm quantize: null pointer");
  if (int rc = check_ld("rmsnorm quantize", x, ldx, K)) return rc;
  return fmt == 0 ? launch_fused<0, 1>(x, ldx, weight, 0, eps, M, K, q, scale, stream)
                  : launch_fused<1, 1>(x, ldx, weight, 0, eps, M, K, q, scale, stream);
}

// SiLU(gate) * up -> per-token quantization.  gate / up bf16 [M, K] with row pitches ldg / ldu (the two halves of a
// fused gate|up projection's output are column slices of one buffer).
extern "C" int ao_silu_mul_quantize_rowwise(const uint16_t* gate, int ldg, const uint16_t* up, int ldu, int M, int K, int fmt,
                                            uint8_t* q, float* scale, void* stream) {
  AO_REQUIRE(M >= 0 && K > 0 && K % 8 == 0 && K <= 48 * 1024, "silu-mul quantize: bad sizes M=%d K=%d (K%%8==0, K<=49152)", M, K);
  AO_REQUIRE(fmt == 0 || fmt == 1, "silu-mul quantize: fmt=%d (0 = int8, 1 = e4m3)", fmt);
  if (M == 0) return AO_OK;
  AO_REQUIRE(gate && up && q && scale, "silu-mul quantize: null pointer");
  if (int rc = check_ld("silu-mul quantize", gate, ldg, K)) return rc;
  if (int rc = check_ld("silu-mul quantize", up, ldu, K)) return rc;
  return fmt == 0 ? launch_fused<0, 2>(gate, ldg, up, ldu, 0.f, M, K, q, scale, stream)
                  : launch_fused<1, 2>(gate, ldg, up, ldu, 0.f, M, K, q, scale, stream);
}
