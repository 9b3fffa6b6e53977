// NVFP4 weight (e2m1 + e4m3 block-16 scales + f32 per-tensor scale) x bf16 activation linear.
//
// Semantics = F.linear(x, NVFP4Tensor.dequantize()) (torchao/prototype/mx_formats/nvfp4_tensor.py:199-231,
// weight-only handler inference_workflow.py:356-400), with an optional per-token activation scale applied in
// the epilogue so that e4m3-rowwise-quantised activations (BASELINE config 5) run on the
// same kernel: their values are exact in bf16, the scale is a row factor.
// tcgen05 has no nvfp4 x bf16/fp8 MMA kind, so the weights are dequantised in-kernel to bf16 (exact: 2
// significant bits x 4 significant bits) and fed from TMEM exactly like the int4 path (ts_gemm.cuh); the
// per-tensor scale is applied in fp32 in the epilogue.
//
// e2m1 -> bf16 without a table: nibble x = (s e1 e0 m) placed at bf16 bits 15|8:6 is the bf16 number
// value(x) * 2^-126 (denormal for e = 0, which bf16 multiplies handle exactly); ONE exact multiply by
// (block_scale * 2^66) gives value(x) * block_scale * 2^-60, and the 2^60 is taken back out of the fp32 accumulator in the
// epilogue (Params::acc_exp2; power-of-two factors commute with every rounding on the way).  bf16(block_scale * 2^66) =
// (byte << 4) + 0x5D00 for the (always normal, >= 2^-6) e4m3 scale bytes.
#include <cuda_bf16.h>
#include <cuda_fp8.h>

#include "common.h"
#include "ptx.cuh"
#include "ts_gemm.cuh"
#include "ts_prefill.cuh"

namespace ao {
namespace nvf4w {

using tsg::KCHUNK;
using tsg::ROWS;
using tsg::W_BYTES;

struct Nvfp4Fmt {
  __device__ static __forceinline__ uint32_t w_tx_bytes(const tsg::Params&) { return W_BYTES + 1024; }
  __device__ static __forceinline__ void issue_w(const CUtensorMap* tm_w, const CUtensorMap* tm_sf, const tsg::Params& p,
                                                 uint8_t* w_dst, uint8_t* aux_dst, uint64_t* bar, int n_tile, int kc,
                                                 uint64_t policy) {
    tma_load_2d(w_dst, tm_w, bar, kc * 64, n_tile * ROWS, policy);  // 128 rows x 64 bytes, 64B swizzle
    // two consecutive blocked scale tiles (128 rows x 4 scales each = 64 k per tile): the blocked scale tensor is a
    // [row block][column block] array of contiguous 512-byte tiles = a 2-D tensor of 128 words x tiles
    tma_load_2d(aux_dst, tm_sf, bar, 0, n_tile * p.aux_col_blocks + kc * 2, policy);
  }
  __device__ static __forceinline__ void prefetch_w(const CUtensorMap* tm_w, const CUtensorMap* tm_sf, const tsg::Params& p, int n_tile, int kc) {
    tma_prefetch_l2_2d(tm_w, kc * 64, n_tile * ROWS);
    tma_prefetch_l2_2d(tm_sf, 0, n_tile * p.aux_col_blocks + kc * 2);
  }
  // one weight row of the chunk: its 64 bytes (k 0..127, even k in the low nibble) and the 8 block scales
  struct Raw {
    uint4 v[4];
    uint32_t sc[2];   // blocked tile h (k 64h..64h+63): four e4m3 scale bytes
  };
  __device__ static __forceinline__ void load_row(const tsg::Params&, uint32_t w_smem, uint32_t aux_smem, int r, Raw& raw) {
    const uint32_t sc_off = (uint32_t)(r & 31) * 16u + (uint32_t)(r >> 5) * 4u;
    raw.sc[0] = tsg::lds32(aux_smem + sc_off);
    raw.sc[1] = tsg::lds32(aux_smem + 512 + sc_off);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t off = (uint32_t)r * 64u + i * 16;
      raw.v[i] = tsg::lds128(w_smem + (off ^ (((off >> 7) & 3) << 4)));  // undo the TMA 64B swizzle
    }
  }
  // depends on one destination register of every ld.shared of load_row (see Int4Fmt::touch)
  __device__ static __forceinline__ uint32_t touch(const Raw& raw) {
    return raw.v[0].x ^ raw.v[1].x ^ raw.v[2].x ^ raw.v[3].x ^ raw.sc[0] ^ raw.sc[1];
  }
  // quarter q = bytes 16q..16q+15 of the row = k 32q..32q+31: out[c] = bf16x2 of k pair (32q + 2c, +1)
  __device__ static __forceinline__ void dequant_quarter(const tsg::Params&, const Raw& raw, int q, uint32_t (&out)[16]) {
    const uint4 v = raw.v[q];
    const uint32_t sc = raw.sc[q >> 1] >> (16 * (q & 1));   // two scale bytes: blocks 2q, 2q + 1
#pragma unroll
    for (int w = 0; w < 4; ++w) {  // word w = k 8w..8w+7 of the quarter; scale block = w/2
      const uint32_t word = w == 0 ? v.x : w == 1 ? v.y : w == 2 ? v.z : v.w;
      const uint32_t sb = (sc >> (8 * (w >> 1))) & 0xFFu;
      // bf16x2 of scale * 2^66: the e4m3 byte re-biased (<< 4, + (127 - 7 + 66) << 7).  2^66 = 2^126 (undoes the
      // denormal placement below) * 2^-60 (ACC_EXP2: taken back out in the epilogue, in fp32) -- one exact multiply per
      // pair instead of two; the product value * scale * 2^-60 has at most 6 significant bits and is far inside the
      // normal bf16 range
      const uint32_t s_bits = ((sb << 4) + 0x5D00u) * 0x00010001u;
      const __nv_bfloat162 s2 = *reinterpret_cast<const __nv_bfloat162*>(&s_bits);
      const uint32_t lo = word & 0x0F0F0F0Fu, hi = (word >> 4) & 0x0F0F0F0Fu;   // even / odd k, one nibble per byte
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // hh = nibble of k = 8w + 2j in bits 3:0, of k + 1 in bits 19:16: one prmt (selector nibble 8|j replicates the
        // sign bit of a byte < 16: zero)
        uint32_t hh;
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(hh) : "r"(lo), "r"(hi), "r"((uint32_t)(((8 | j) << 12) | ((4 + j) << 8) | ((8 | j) << 4) | j)));
        const uint32_t bits = (hh * 0x1040u) & 0x81C081C0u;           // sign | e1 e0 m at bf16 bits 15 | 8:6 = value * 2^-126
        __nv_bfloat162 x = *reinterpret_cast<const __nv_bfloat162*>(&bits);
        x = __hmul2(x, s2);
        out[4 * w + j] = *reinterpret_cast<uint32_t*>(&x);
      }
    }
  }
  static constexpr int ACC_EXP2 = 60;   // the accumulators hold the result * 2^-60 (Params::acc_exp2)
};

// N_MMA = tokens per tile of the decode kernel (16 .. 128), or 0 = the prefill-shaped kernel (ts_prefill.cuh, 256 tokens)
template <int N_MMA>
static int launch_tc(const uint16_t* x, int ldx, const float* x_scale, int M, int K, const uint8_t* wq, const uint8_t* w_sf,
                     const float* b_pts, int b_pts_per_row, int N, const uint16_t* bias, uint16_t* y, void* ws,
                     size_t ws_bytes, cudaStream_t stream) {
  constexpr bool PREFILL = N_MMA == 0;
  constexpr int TOK = PREFILL ? tsp::N_TOK : N_MMA;
  CUtensorMap tm_w, tm_x;
  {
    const uint64_t dims[2] = {(uint64_t)K / 2, (uint64_t)N};
    const uint64_t str[1] = {(uint64_t)K / 2};
    const uint32_t box[2] = {64, 128};
    int rc = make_tmap(&tm_w, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, wq, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
  }
  CUtensorMap tm_sf;
  {
    const uint64_t dims[2] = {128, (uint64_t)ceil_div(N, ROWS) * (uint64_t)ceil_div(K / 16, 4)};
    const uint64_t str[1] = {512};
    const uint32_t box[2] = {128, 2};
    int rc = make_tmap(&tm_sf, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, w_sf, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    const uint64_t str[1] = {(uint64_t)ldx * 2};
    const uint32_t box[2] = {64, (uint32_t)TOK};
    int rc = make_tmap(&tm_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  tsg::Params p{};
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.row_scale = x_scale;
  p.out_scale = b_pts;
  p.out_scale_per_row = b_pts_per_row;
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.aux_base = w_sf;
  p.acc_exp2 = Nvfp4Fmt::ACC_EXP2;
  p.aux_col_blocks = ceil_div(K / 16, 4);
  p.M = M; p.N = N; p.N_out = N; p.K = K; p.group_size = 16;
  p.n_tiles = ceil_div(N, ROWS);
  p.m_blocks = ceil_div(M, TOK);
  p.KT = K / 128;
  int grid = 0;
  if constexpr (PREFILL) {
    if (int rc = tsp::plan(p, ws, ws_bytes, "nvfp4 weight linear (prefill)", &grid)) return rc;
    auto kern = tsp::ts_prefill_kernel<Nvfp4Fmt>;
    AO_CUDA_CHECK(ensure_dynamic_smem(reinterpret_cast<const void*>(kern), tsp::SMEM_BYTES));
    AO_CUDA_CHECK(launch(kern, dim3(grid), dim3(tsp::NUM_THREADS), tsp::SMEM_BYTES, stream, pdl_enabled(), tm_w, tm_sf, tm_x, p));
  } else {
    using C = tsg::Cfg<(PREFILL ? 128 : N_MMA)>;
    if (int rc = tsg::plan<(PREFILL ? 128 : N_MMA)>(p, ws, ws_bytes, "nvfp4 weight linear", &grid)) return rc;
    p.timeline = nullptr;
    auto kern = tsg::ts_gemm_kernel<Nvfp4Fmt, (PREFILL ? 128 : N_MMA)>;
    AO_CUDA_CHECK(ensure_dynamic_smem(reinterpret_cast<const void*>(kern), C::SMEM_BYTES));
    AO_CUDA_CHECK(launch(kern, dim3(grid), dim3(tsg::NUM_THREADS), C::SMEM_BYTES, stream, pdl_enabled(), tm_w, tm_sf, tm_x, p));
  }
  return AO_OK;
}

// per-token e4m3 "fake quantisation": x -> bf16(e4m3(x / s)) and s = f32(bf16(amax/448)); the bf16 values
// are exactly the e4m3 codes Float8Tensor.from_hp(x, PerRow()) would store (quant_primitives.py:2172-2287).
__global__ void __launch_bounds__(256) fp8_fakequant_rowwise_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int K,
                                                                    __nv_bfloat16* __restrict__ xq,
                                                                    float* __restrict__ scale) {
  __shared__ float sh[8];
  // PDL: let the linear that consumes this output become resident and start its weight stream now; our own input may be
  // the previous kernel's output, so wait for it before the first read
  pdl_launch_dependents();
  pdl_wait();
  const int m = blockIdx.x;
  const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)m * ldx);
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
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = amax;
  __syncthreads();
  amax = sh[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) amax = fmaxf(amax, sh[i]);
  const float s = __bfloat162float(__float2bfloat16_rn(amax / 448.0f));
  if (threadIdx.x == 0) scale[m] = s;
  uint4* qr = reinterpret_cast<uint4*>(xq + (size_t)m * K);
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    const uint4 v = xr[i];
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
    __nv_bfloat16 o[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(h[j]);
      float a = fminf(fmaxf(f.x / s, -448.f), 448.f), b = fminf(fmaxf(f.y / s, -448.f), 448.f);
      if (s == 0.f) { a = 0.f; b = 0.f; }  // all-zero row: the reference yields NaN (0/0); we keep zeros
      const __nv_fp8_storage_t qa = __nv_cvt_float_to_fp8(a, __NV_SATFINITE, __NV_E4M3);
      const __nv_fp8_storage_t qb = __nv_cvt_float_to_fp8(b, __NV_SATFINITE, __NV_E4M3);
      o[2 * j] = __float2bfloat16_rn(__half2float(__half(__nv_cvt_fp8_to_halfraw(qa, __NV_E4M3))));
      o[2 * j + 1] = __float2bfloat16_rn(__half2float(__half(__nv_cvt_fp8_to_halfraw(qb, __NV_E4M3))));
    }
    qr[i] = *reinterpret_cast<const uint4*>(o);
  }
}

}  // namespace nvf4w
}  // namespace ao

using namespace ao;

extern "C" int ao_fp8_fakequant_rowwise_ld(const uint16_t* x, int ldx, int M, int K, uint16_t* xq_bf16, float* scale,
                                           void* stream) {
  AO_REQUIRE(M >= 0 && K > 0 && K % 8 == 0, "fp8 fakequant: bad sizes M=%d K=%d", M, K);
  if (M == 0) return AO_OK;
  AO_REQUIRE(x && xq_bf16 && scale, "fp8 fakequant: null pointer");
  AO_REQUIRE(ldx >= K && ldx % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0,
             "fp8 fakequant: ldx=%d must be >= K=%d, a multiple of 8, x 16-byte aligned", ldx, K);
  AO_CUDA_CHECK(ao::launch(nvf4w::fp8_fakequant_rowwise_kernel, dim3(M), dim3(256), 0,
                           reinterpret_cast<cudaStream_t>(stream), pdl_enabled(), reinterpret_cast<const __nv_bfloat16*>(x), ldx, K,
                           reinterpret_cast<__nv_bfloat16*>(xq_bf16), scale));
  return AO_OK;
}
extern "C" int ao_fp8_fakequant_rowwise(const uint16_t* x, int M, int K, uint16_t* xq_bf16, float* scale, void* stream) {
  return ao_fp8_fakequant_rowwise_ld(x, K, M, K, xq_bf16, scale, stream);
}

extern "C" int ao_nvfp4_weight_linear_ex(const uint16_t* x, int ldx, const float* x_scale, int M, int K, const uint8_t* wq,
                                         const uint8_t* w_scale_blocked, const float* b_pts, int b_pts_per_row, int N,
                                         const uint16_t* bias, uint16_t* y, void* workspace, size_t workspace_bytes,
                                         void* stream) {
  AO_REQUIRE(M >= 0 && K > 0 && N > 0, "nvfp4 weight linear: bad sizes M=%d K=%d N=%d", M, K, N);
  AO_REQUIRE(K % 128 == 0, "nvfp4 weight linear: K=%d must be a multiple of 128", K);
  AO_REQUIRE(N % 16 == 0, "nvfp4 weight linear: N=%d must be a multiple of 16 (inference_workflow.py:248-251)", N);
  AO_REQUIRE(ldx >= K && ldx % 8 == 0, "nvfp4 weight linear: ldx=%d must be >= K=%d and a multiple of 8", ldx, K);
  if (M == 0) return AO_OK;
  AO_REQUIRE(x && wq && w_scale_blocked && y, "nvfp4 weight linear: null pointer");
  AO_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "nvfp4 weight linear: x must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (M <= 16) return nvf4w::launch_tc<16>(x, ldx, x_scale, M, K, wq, w_scale_blocked, b_pts, b_pts_per_row, N, bias, y, workspace, workspace_bytes, st);
  if (M <= 32) return nvf4w::launch_tc<32>(x, ldx, x_scale, M, K, wq, w_scale_blocked, b_pts, b_pts_per_row, N, bias, y, workspace, workspace_bytes, st);
  if (M <= 64) return nvf4w::launch_tc<64>(x, ldx, x_scale, M, K, wq, w_scale_blocked, b_pts, b_pts_per_row, N, bias, y, workspace, workspace_bytes, st);
  if (!tsp::worth_it(M, N, K))
    return nvf4w::launch_tc<128>(x, ldx, x_scale, M, K, wq, w_scale_blocked, b_pts, b_pts_per_row, N, bias, y, workspace, workspace_bytes, st);
  return nvf4w::launch_tc<0>(x, ldx, x_scale, M, K, wq, w_scale_blocked, b_pts, b_pts_per_row, N, bias, y, workspace, workspace_bytes, st);
}

extern "C" int ao_nvfp4_weight_linear(const uint16_t* x, const float* x_scale, int M, int K, const uint8_t* wq,
                                      const uint8_t* w_scale_blocked, const float* b_pts, int N,
                                      const uint16_t* bias, uint16_t* y, void* workspace, size_t workspace_bytes,
                                      void* stream) {
  return ao_nvfp4_weight_linear_ex(x, K, x_scale, M, K, wq, w_scale_blocked, b_pts, 0, N, bias, y, workspace, workspace_bytes,
                                   stream);
}
