// int4 weight-only linear on the tinygemm "tile_packed_to_4d" format, sm_100a.
//
//   Y[M,N] = X[M,K] * W^[N,K]^T (+bias),  W^ = bf16((q-8)*s + z)   (group-wise s,z)
//
// Replaces aten._weight_int4pack_mm as called from the reference handler
// (torchao/quantization/quantize_/workflows/int4/int4_tile_packed_to_4d_tensor.py:243-299).
// The GEMM itself is the persistent tcgen05 TS-mode kernel of ts_gemm.cuh; this file supplies the
// int4 format policy (how a 128-row x 128-k chunk is fetched and turned into bf16) and the launcher.
//
// qdata layout (int32 [N/8][K/128][32][4], inner_k_tiles = 8), word `wd` of lane `t`:
//   row n = 8*n8 + t/4;  k0 = 128*ko + 32*wd + 2*(t%4);
//   bits [4e,4e+4)   = q[n, k0 + 8e]      e = 0..3
//   bits [16+4e, ..) = q[n, k0 + 8e + 1]
// so one row's 128 k of a k-tile are the 64 contiguous bytes of lanes 4*(n%8)..+3, a 128-row chunk is
// 16 runs of 512 B, fetched by ONE 3-D TMA box {32 words, 4 row-pairs, 16 n8-tiles} with 128-byte
// swizzle -- which makes the per-thread 16-byte ld.shared.v4 of "its" row bank-conflict free (Int4Fmt::load_row).
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"
#include "ts_gemm.cuh"
#include "ts_prefill.cuh"

namespace ao {
namespace int4k {

using tsg::KCHUNK;
using tsg::ROWS;
using tsg::W_BYTES;

// (128+q) bf16x2 bits -> bf16x2 of fma(q-8, s, z), single rounding, = oracle W^.
__device__ __forceinline__ uint32_t deq_pair(uint32_t magic_bits, __nv_bfloat162 s2, __nv_bfloat162 z2) {
  const __nv_bfloat162 c136 = __floats2bfloat162_rn(136.f, 136.f);
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&magic_bits);
  v = __hsub2(v, c136);    // exact: (128+q) - 136 = q - 8
  v = __hfma2(v, s2, z2);  // bf16(fma(q-8, s, z))
  return *reinterpret_cast<uint32_t*>(&v);
}

struct Int4Fmt {
  __device__ static __forceinline__ uint32_t w_tx_bytes(const tsg::Params& p) {
    const int gpc = p.group_size <= KCHUNK ? KCHUNK / p.group_size : 1;
    return W_BYTES + gpc * 512;
  }
  __device__ static __forceinline__ void issue_w(const CUtensorMap* tm_w, const CUtensorMap* tm_sz,
                                                 const tsg::Params& p, uint8_t* w_dst, uint8_t* aux_dst,
                                                 uint64_t* bar, int n_tile, int kc, uint64_t policy) {
    tma_load_3d(w_dst, tm_w, bar, 0, 4 * kc, n_tile * (ROWS / 8), policy);
    tma_load_2d(aux_dst, tm_sz, bar, n_tile * ROWS, (kc * KCHUNK) / p.group_size, policy);
  }
  // One weight row of the chunk in registers: the row's 64 bytes are the four 16-byte lane words of its n8 group
  // (tinygemm word wd of lane word i holds k = 32wd + 2i + 8e + {0,1}), plus the (s, z) pair of each 32-k word.
  struct Raw {
    uint4 v[4];
    uint32_t sz[4];
  };
  __device__ static __forceinline__ void prefetch_w(const CUtensorMap* tm_w, const CUtensorMap* tm_sz, const tsg::Params& p, int n_tile, int kc) {
    tma_prefetch_l2_3d(tm_w, 0, 4 * kc, n_tile * (ROWS / 8));
    tma_prefetch_l2_2d(tm_sz, n_tile * ROWS, (kc * KCHUNK) / p.group_size);
  }
  // thread r (= TMEM lane = weight row of the tile): 4 x ld.shared.v4 through the TMA 128-byte swizzle -- per
  // quarter-warp the eight rows hit eight different 16-byte bank groups: conflict-free (the half-row ld.shared.v2
  // of round 1 was 2-way conflicted: profiles/r01_int4_final_ncu_full.csv) -- and 4 x ld.shared.b32 of (s, z)
  __device__ static __forceinline__ void load_row(const tsg::Params& p, uint32_t w_smem, uint32_t aux_smem, int r, Raw& raw) {
    const int gshift = p.group_size == 32 ? 0 : (p.group_size == 64 ? 1 : 2);  // 32-k word -> group of the chunk
#pragma unroll
    for (int w = 0; w < 4; ++w) raw.sz[w] = tsg::lds32(aux_smem + r * 4 + (w >> gshift) * 512);
    const uint32_t row_off = (uint32_t)(r >> 3) * 512u + (uint32_t)(r & 7) * 64u;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t off = row_off + i * 16;
      raw.v[i] = tsg::lds128(w_smem + (off ^ (((off >> 7) & 7) << 4)));  // undo the TMA 128B swizzle
    }
  }
  // A value that depends on one destination register of EVERY ld.shared of load_row.  The kernels fold it into the
  // address of the mbarrier arrive that hands the stage back to the TMA producers (through a comparison that is never
  // true at run time but that ptxas cannot decide), so the arrive cannot issue before the loads have returned.  Without a
  // data dependence it does: an empty asm with "r" inputs emits nothing, an unused xor chain is removed by ptxas, the
  // arrive overtakes the loads in flight and the refill races with them (round 2: 2 of 6 runs of the nvfp4-weight prefill
  // test off by a few values, profiles/r02_call_s.log).
  __device__ static __forceinline__ uint32_t touch(const Raw& raw) {
    return raw.v[0].x ^ raw.v[1].x ^ raw.v[2].x ^ raw.v[3].x ^ raw.sz[0] ^ raw.sz[1] ^ raw.sz[2] ^ raw.sz[3];
  }
  // quarter q = the 32 k of tinygemm word q: out[c] = bf16x2 of k pair (32q + 2c, 32q + 2c + 1), c = i + 4e
  __device__ static __forceinline__ void dequant_quarter(const tsg::Params&, const Raw& raw, int q, uint32_t (&out)[16]) {
    uint32_t magic = 0x43004300u;
    asm volatile("" : "+r"(magic));  // keep it in a register
    const uint32_t sz = raw.sz[q];
    const uint32_t s_bits = __byte_perm(sz, sz, 0x1010);
    const uint32_t z_bits = __byte_perm(sz, sz, 0x3232);
    const __nv_bfloat162 s2 = *reinterpret_cast<const __nv_bfloat162*>(&s_bits);
    const __nv_bfloat162 z2 = *reinterpret_cast<const __nv_bfloat162*>(&z_bits);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t word = q == 0 ? raw.v[i].x : q == 1 ? raw.v[i].y : q == 2 ? raw.v[i].z : raw.v[i].w;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        // bf16x2 of 128+q = ((word >> 4e) & 0x000F000F) | 0x43004300 as ONE lop3 (C source compiles to two LOP3
        // because both constants want the immediate slot)
        uint32_t m;
        asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(m) : "r"(word >> (4 * e)), "r"(0x000F000Fu), "r"(magic));
        out[i + 4 * e] = deq_pair(m, s2, z2);   // k pair (32q + 2i + 8e, +1)
      }
    }
  }
};

// ---------------------------------------------------------------------------------------
// Reference-grade CUDA-core kernel (impl = 2): one warp per weight row, lanes stride over
// k-tiles, fp32 accumulation.  Slow; used as an on-device cross-check and for odd shapes.
template <int MT>
__global__ void int4_linear_simple_kernel(const __nv_bfloat16* __restrict__ x,
                                          const int32_t* __restrict__ qdata,
                                          const __nv_bfloat16* __restrict__ sz,
                                          const __nv_bfloat16* __restrict__ bias,
                                          __nv_bfloat16* __restrict__ y, int M, int N, int N_out,
                                          int K, int g, int ldx) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  const int m0 = blockIdx.y * MT;
  if (n >= N) return;
  const int KT = K / 128;
  float acc[MT];
#pragma unroll
  for (int m = 0; m < MT; ++m) acc[m] = 0.f;
  for (int ko = lane; ko < KT; ko += 32) {
    const uint4* wp = reinterpret_cast<const uint4*>(qdata + (((size_t)(n >> 3) * KT + ko) * 32 + (n & 7) * 4) * 4);
    for (int tq = 0; tq < 4; ++tq) {
      const uint4 v = wp[tq];
      const uint32_t words[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int wd = 0; wd < 4; ++wd) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int k = ko * 128 + 32 * wd + 2 * tq + 8 * e + h;
            const int q = (words[wd] >> (4 * e + 16 * h)) & 15;
            const size_t gi = ((size_t)(k / g) * N + n) * 2;
            const __nv_bfloat16 wv = __hfma(__int2bfloat16_rn(q - 8), sz[gi], sz[gi + 1]);
            const float wf = __bfloat162float(wv);
#pragma unroll
            for (int m = 0; m < MT; ++m)
              if (m0 + m < M) acc[m] += __bfloat162float(x[(size_t)(m0 + m) * ldx + k]) * wf;
          }
        }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    float v = acc[m];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && m0 + m < M && n < N_out)
      y[(size_t)(m0 + m) * N_out + n] = __float2bfloat16_rn(v + (bias ? __bfloat162float(bias[n]) : 0.f));
  }
}


// tensor maps of the packed weights ({32 words, 4 row-pairs, 16 n8-tiles} box, 128-byte swizzle), the (scale, zero)
// pairs and the activations (box = 64 k x `x_rows` tokens)
static int make_maps(const uint16_t* x, int ldx, int M, int K, const int32_t* qdata, const uint16_t* sz, int g, int N,
                     int x_rows, CUtensorMap* tm_w, CUtensorMap* tm_sz, CUtensorMap* tm_x) {
  const int KT = K / 128;
  {
    const uint64_t dims[3] = {32, (uint64_t)4 * KT, (uint64_t)N / 8};
    const uint64_t str[2] = {128, (uint64_t)KT * 512};
    const uint32_t box[3] = {32, 4, 16};
    int rc = make_tmap(tm_w, CU_TENSOR_MAP_DATA_TYPE_INT32, 3, qdata, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  const int gpc = g <= 128 ? 128 / g : 1;
  {
    const uint64_t dims[2] = {(uint64_t)N, (uint64_t)K / g};
    const uint64_t str[1] = {(uint64_t)N * 4};
    const uint32_t box[2] = {128, (uint32_t)gpc};
    int rc = make_tmap(tm_sz, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, sz, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    const uint64_t str[1] = {(uint64_t)ldx * 2};
    const uint32_t box[2] = {64, (uint32_t)x_rows};
    int rc = make_tmap(tm_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  return AO_OK;
}

template <int N_MMA, int DBUF = 2>
static int launch_tc(const uint16_t* x, int ldx, int M, int K, const int32_t* qdata, const uint16_t* sz, int g, int N,
                     const uint16_t* bias, uint16_t* y, int N_out, void* ws, size_t ws_bytes,
                     cudaStream_t stream) {
  using C = tsg::Cfg<N_MMA, DBUF>;
  const int KT = K / 128;
  CUtensorMap tm_w, tm_sz, tm_x;
  if (int rc = make_maps(x, ldx, M, K, qdata, sz, g, N, N_MMA, &tm_w, &tm_sz, &tm_x)) return rc;
  tsg::Params p{};
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.M = M; p.N = N; p.N_out = N_out; p.K = K; p.group_size = g;
  p.n_tiles = ceil_div(N_out, ROWS);
  p.m_blocks = ceil_div(M, N_MMA);
  p.KT = KT;
  int grid = 0;
  if (int rc = tsg::plan<N_MMA>(p, ws, ws_bytes, "int4 linear", &grid)) return rc;
  // bring-up timeline: two slots (consecutive launches alternate) of 100 CTAs x 16 stamps + 8 chunks x 8 fine stamps at workspace + 20 MiB
  static unsigned tl_launch = 0;
  p.timeline = timeline_enabled() ? reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(ws) + ((size_t)20 << 20) +
                                                                         (size_t)(tl_launch++ & 1) * (100 * 16 + 64) * 8)
                                  : nullptr;
  auto kern = (p.timeline && N_MMA <= 32) ? tsg::ts_gemm_kernel<Int4Fmt, (N_MMA <= 32 ? N_MMA : 16), true, DBUF>
                                           : tsg::ts_gemm_kernel<Int4Fmt, N_MMA, false, DBUF>;
  AO_CUDA_CHECK(ensure_dynamic_smem(reinterpret_cast<const void*>(kern), C::SMEM_BYTES));
  AO_CUDA_CHECK(launch(kern, dim3(grid), dim3(tsg::NUM_THREADS), C::SMEM_BYTES, stream, pdl_enabled(), tm_w, tm_sz,
                       tm_x, p));
  return AO_OK;
}

// M > 128 tokens: the prefill-shaped kernel (ts_prefill.cuh), weights dequantised once per 256 tokens
static int launch_prefill(const uint16_t* x, int ldx, int M, int K, const int32_t* qdata, const uint16_t* sz, int g, int N,
                          const uint16_t* bias, uint16_t* y, int N_out, void* ws, size_t ws_bytes, cudaStream_t stream) {
  CUtensorMap tm_w, tm_sz, tm_x;
  if (int rc = make_maps(x, ldx, M, K, qdata, sz, g, N, tsp::N_TOK, &tm_w, &tm_sz, &tm_x)) return rc;
  tsg::Params p{};
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.M = M; p.N = N; p.N_out = N_out; p.K = K; p.group_size = g;
  p.n_tiles = ceil_div(N_out, ROWS);
  p.m_blocks = ceil_div(M, tsp::N_TOK);
  p.KT = K / 128;
  int grid = 0;
  if (int rc = tsp::plan(p, ws, ws_bytes, "int4 linear (prefill)", &grid)) return rc;
  auto kern = tsp::ts_prefill_kernel<Int4Fmt>;
  AO_CUDA_CHECK(ensure_dynamic_smem(reinterpret_cast<const void*>(kern), tsp::SMEM_BYTES));
  AO_CUDA_CHECK(launch(kern, dim3(grid), dim3(tsp::NUM_THREADS), tsp::SMEM_BYTES, stream, pdl_enabled(), tm_w, tm_sz,
                       tm_x, p));
  return AO_OK;
}

}  // namespace int4k
}  // namespace ao

extern "C" int ao_int4_tilepacked_linear_strided(const uint16_t* x, int ldx, int M, int K, const int32_t* qdata,
                                                 const uint16_t* scale_and_zero, int group_size, int N,
                                                 const uint16_t* bias, uint16_t* y, int N_out,
                                                 void* workspace, size_t workspace_bytes, int impl,
                                                 void* stream) {
  using namespace ao;
  AO_REQUIRE(M >= 0 && K > 0 && N > 0, "int4 linear: bad sizes M=%d K=%d N=%d", M, K, N);
  AO_REQUIRE(K % 1024 == 0, "int4 linear: K=%d must be a multiple of 1024 (format pads K)", K);
  AO_REQUIRE(N % 8 == 0, "int4 linear: N=%d must be a multiple of 8", N);
  AO_REQUIRE(N_out > 0 && N_out <= N, "int4 linear: N_out=%d out of range (N=%d)", N_out, N);
  AO_REQUIRE(group_size == 32 || group_size == 64 || group_size == 128 || group_size == 256,
             "int4 linear: group_size=%d not in {32,64,128,256}", group_size);
  AO_REQUIRE(ldx >= K && ldx % 8 == 0, "int4 linear: ldx=%d must be >= K=%d and a multiple of 8", ldx, K);
  if (M == 0) return AO_OK;
  AO_REQUIRE(x && qdata && scale_and_zero && y, "int4 linear: null pointer");
  AO_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "int4 linear: x must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (impl == 2) {
    constexpr int MT = 8;
    dim3 grid(ceil_div(N, 8), ceil_div(M, MT));
    AO_CUDA_CHECK(launch(int4k::int4_linear_simple_kernel<MT>, grid, dim3(256), 0, st, false,
                         reinterpret_cast<const __nv_bfloat16*>(x), qdata,
                         reinterpret_cast<const __nv_bfloat16*>(scale_and_zero),
                         reinterpret_cast<const __nv_bfloat16*>(bias),
                         reinterpret_cast<__nv_bfloat16*>(y), M, N, N_out, K, group_size, ldx));
    return AO_OK;
  }
  if (M <= 16)
    return int4k::launch_tc<16>(x, ldx, M, K, qdata, scale_and_zero, group_size, N, bias, y, N_out, workspace,
                                workspace_bytes, st);
  if (M <= 32)
    return int4k::launch_tc<32>(x, ldx, M, K, qdata, scale_and_zero, group_size, N, bias, y, N_out, workspace,
                                workspace_bytes, st);
  if (M <= 64)
    return int4k::launch_tc<64>(x, ldx, M, K, qdata, scale_and_zero, group_size, N, bias, y, N_out, workspace,
                                workspace_bytes, st);
  if (!tsp::worth_it(M, N_out, K))
    return int4k::launch_tc<128>(x, ldx, M, K, qdata, scale_and_zero, group_size, N, bias, y, N_out, workspace,
                                 workspace_bytes, st);
  return int4k::launch_prefill(x, ldx, M, K, qdata, scale_and_zero, group_size, N, bias, y, N_out, workspace,
                               workspace_bytes, st);
}

extern "C" int ao_int4_tilepacked_linear(const uint16_t* x, int M, int K, const int32_t* qdata,
                                         const uint16_t* scale_and_zero, int group_size, int N,
                                         const uint16_t* bias, uint16_t* y, int N_out,
                                         void* workspace, size_t workspace_bytes, int impl,
                                         void* stream) {
  return ao_int4_tilepacked_linear_strided(x, K, M, K, qdata, scale_and_zero, group_size, N, bias, y, N_out, workspace,
                                           workspace_bytes, impl, stream);
}
