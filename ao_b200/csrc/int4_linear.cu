// int4 weight-only linear on the tinygemm "tile_packed_to_4d" format, sm_100a.
//
//   Y[M,N] = X[M,K] * W^[N,K]^T (+bias),  W^ = bf16((q-8)*s + z)   (group-wise s,z)
//
// Replaces aten._weight_int4pack_mm as called from the reference handler
// (torchao/quantization/quantize_/workflows/int4/int4_tile_packed_to_4d_tensor.py:243-299).
//
// Design (decode-shaped, M <= 128 tokens per block): swap-AB tcgen05 GEMM.
//   * the 128 weight rows of a CTA are the UMMA M dimension; tokens are the UMMA N dimension
//   * TMA (SWIZZLE_128B) streams 8 KiB packed-weight chunks (128 rows x 128 k) + the (s,z)
//     rows + the activation tile into a deep smem ring
//   * 8 dequant warps (two warpgroups alternating chunks) unpack nibbles with the bf16
//     magic-number trick, apply fma(q-8, s, z) in bf16x2 (bit-identical to the oracle's W^)
//     and write the bf16 A operand straight into TENSOR MEMORY with tcgen05.st
//   * one thread issues tcgen05.mma.kind::f16 with A from TMEM, B (activations) from smem,
//     fp32 accumulator in TMEM;  tcgen05.commit recycles the TMEM / smem stages
//   * split-K over CTAs with a deterministic last-CTA reduction through an fp32 workspace
//   * PDL: weights/scales are prefetched before griddepcontrol.wait, only activations wait
//
// qdata layout (int32 [N/8][K/128][32][4], inner_k_tiles = 8), word `wd` of lane `t`:
//   row n = 8*n8 + t/4;  k0 = 128*ko + 32*wd + 2*(t%4);
//   bits [4e,4e+4)   = q[n, k0 + 8e]      e = 0..3
//   bits [16+4e, ..) = q[n, k0 + 8e + 1]
// so one row's 128 k of a k-tile are the 64 contiguous bytes of lanes 4*(n%8)..+3.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace ao {
namespace int4k {

constexpr int ROWS = 128;               // weight rows per CTA (UMMA M)
constexpr int KCHUNK = 128;             // k per pipeline stage
constexpr int W_BYTES = ROWS * KCHUNK / 2;  // 8192
constexpr int SZ_BYTES = 2048;          // up to 4 groups x 128 rows x (s,z)
constexpr int A_STAGES = 4;             // TMEM A-operand stages (64 columns each)
constexpr int A_COLS = 64;
constexpr int TMA_WARP = 8;
constexpr int MMA_WARP = 9;
constexpr int NUM_THREADS = 320;
constexpr int TMEM_COLS = 512;
constexpr int MAX_SPLITS = 16;

template <int N_MMA>
struct Cfg {
  static constexpr int X_BYTES = 2 * N_MMA * 128;  // two 64-k swizzle atoms
  static constexpr int STAGE_BYTES = W_BYTES + X_BYTES + SZ_BYTES;
  static constexpr int MAX_STAGES = (196 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = MAX_STAGES > 12 ? 12 : MAX_STAGES;
  static constexpr int D_COL = 0;              // accumulator columns [0, N_MMA)
  static constexpr int A_COL0 = 128;           // A stages at columns [128, 128 + 4*64)
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*align*/ + 1024 /*bars*/;
};

struct Params {
  const __nv_bfloat16* bias;  // [N_out] or null
  __nv_bfloat16* y;           // [M, N_out]
  float* ws_partial;
  unsigned int* ws_sem;
  int M, N, N_out, K, group_size;
  int splits;  // gridDim.y
};

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// (128+q) bf16x2 bits -> bf16x2 of fma(q-8, s, z), single rounding, = oracle W^.
__device__ __forceinline__ uint32_t deq_pair(uint32_t magic_bits, __nv_bfloat162 s2,
                                             __nv_bfloat162 z2) {
  const __nv_bfloat162 c136 = __floats2bfloat162_rn(136.f, 136.f);
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&magic_bits);
  v = __hsub2(v, c136);           // exact: (128+q) - 136 = q - 8
  v = __hfma2(v, s2, z2);         // bf16(fma(q-8, s, z))
  return *reinterpret_cast<uint32_t*>(&v);
}

// DBG: 0 = production; 1 = no dequant math (raw words to TMEM); 2 = math but no TMEM store;
//      3 = dequant warps only recycle the stages (pure TMA/barrier pipeline); 4 = 3 + no MMAs;
//      5 = 3 + MMAs round-robin over 4 accumulators; 6 = 3 + one MMA per chunk.  Bring-up only.
template <int N_MMA, int DBG>
__global__ void __launch_bounds__(NUM_THREADS, 1)
int4_linear_tc_kernel(const __grid_constant__ CUtensorMap tm_w,
                      const __grid_constant__ CUtensorMap tm_sz,
                      const __grid_constant__ CUtensorMap tm_x, const Params p) {
  using C = Cfg<N_MMA>;
  constexpr int S = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * C::STAGE_BYTES);
  uint64_t* wfull = bars;                 // [S]  TMA -> dequant   (weights + scales)
  uint64_t* xfull = bars + S;             // [S]  TMA -> MMA       (activations)
  uint64_t* sempty = bars + 2 * S;        // [S]  dequant(4 warps) + MMA commit -> TMA
  uint64_t* afull = bars + 3 * S;         // [A_STAGES] dequant -> MMA
  uint64_t* aempty = afull + A_STAGES;    // [A_STAGES] MMA commit -> dequant
  uint64_t* dfull = aempty + A_STAGES;    // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dfull + 1);
  uint32_t* flag_slot = tmem_slot + 1;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x;
  const int split = blockIdx.y;
  const int m_blk = blockIdx.z;
  const int n0 = n_tile * ROWS;
  const int m0 = m_blk * N_MMA;

  // balanced chunk range of this split
  const int total_chunks = p.K / KCHUNK;
  const int c_begin = (int)(((long long)total_chunks * split) / p.splits);
  const int c_end = (int)(((long long)total_chunks * (split + 1)) / p.splits);
  const int nchunks = c_end - c_begin;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&wfull[i], 1);
      mbar_init(&xfull[i], 1);
      mbar_init(&sempty[i], 5);
    }
    for (int i = 0; i < A_STAGES; ++i) {
      mbar_init(&afull[i], 4);
      mbar_init(&aempty[i], 1);
    }
    mbar_init(dfull, 1);
    fence_barrier_init();
  }
  if (warp == TMA_WARP && lane == 0) {
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_sz);
    tma_prefetch_desc(&tm_x);
  }
  if (warp == MMA_WARP) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // let the next kernel in the stream start its own weight prefetch as early as possible
  pdl_launch_dependents();

  const int gpc = p.group_size <= KCHUNK ? KCHUNK / p.group_size : 1;  // (s,z) rows per chunk

  if (warp == TMA_WARP) {
    if (lane == 0) {
      const uint64_t pol_w = policy_evict_first();
      const uint64_t pol_x = policy_evict_last();
      const uint32_t w_tx = W_BYTES + gpc * 512;
      auto issue_w = [&](int c) {
        const int s = c % S;
        uint8_t* st = smem + (size_t)s * C::STAGE_BYTES;
        const int kc = c_begin + c;
        mbar_expect_tx(&wfull[s], w_tx);
        tma_load_3d(st, &tm_w, &wfull[s], 0, 4 * kc, n0 / 8, pol_w);
        tma_load_2d(st + W_BYTES + C::X_BYTES, &tm_sz, &wfull[s], n0,
                    (kc * KCHUNK) / p.group_size, pol_w);
      };
      auto issue_x = [&](int c) {
        const int s = c % S;
        uint8_t* st = smem + (size_t)s * C::STAGE_BYTES + W_BYTES;
        const int kc = c_begin + c;
        mbar_expect_tx(&xfull[s], C::X_BYTES);
        tma_load_2d(st, &tm_x, &xfull[s], kc * KCHUNK, m0, pol_x);
        tma_load_2d(st + N_MMA * 128, &tm_x, &xfull[s], kc * KCHUNK + 64, m0, pol_x);
      };
      const int pre = nchunks < S ? nchunks : S;
      for (int c = 0; c < pre; ++c) issue_w(c);   // weights do not depend on the previous kernel
      pdl_wait();                                  // activations do
      for (int c = 0; c < pre; ++c) issue_x(c);
      for (int c = S; c < nchunks; ++c) {
        mbar_wait(&sempty[c % S], ((c / S) & 1) ^ 1);
        issue_w(c);
        issue_x(c);
      }
    }
  } else if (warp == MMA_WARP) {
    constexpr uint32_t idesc = make_idesc(1 /*f32*/, 1 /*bf16*/, 1 /*bf16*/, ROWS, N_MMA);
    for (int c = 0; c < nchunks; ++c) {
      const int s = c % S, t = c % A_STAGES;
      mbar_wait(&xfull[s], (c / S) & 1);
      mbar_wait(&afull[t], (c / A_STAGES) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t xb = smem_u32(smem + (size_t)s * C::STAGE_BYTES + W_BYTES);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          if (DBG == 4) break;
          if (DBG == 6 && kk > 0) break;
          const uint64_t bdesc = umma_desc_k_sw128(xb + (kk >> 2) * (N_MMA * 128) + (kk & 3) * 32);
          const uint32_t a_t = tmem_base + C::A_COL0 + t * A_COLS + kk * 8;
          const uint32_t d_t = tmem_base + C::D_COL + (DBG == 5 ? (kk & 3) * N_MMA : 0);
          mma_ts_f16(d_t, a_t, bdesc, idesc, (c > 0 || kk > (DBG == 5 ? 3 : 0)) ? 1u : 0u);
        }
        tc_commit(&aempty[t]);
        tc_commit(&sempty[s]);
        if (c == nchunks - 1) tc_commit(dfull);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ dequant warps
    const int wg = warp >> 2;                 // warpgroup 0/1: even / odd chunks
    const int q4 = warp & 3;                  // TMEM lane quarter
    const int r = q4 * 32 + lane;             // weight row within the tile == TMEM lane
    const uint32_t row_off = (uint32_t)(r >> 3) * 512u + (uint32_t)(r & 7) * 64u;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const int gshift = p.group_size == 32 ? 0 : (p.group_size == 64 ? 1 : 2);  // word -> group

    for (int c = wg; c < nchunks; c += 2) {
      const int s = c % S, t = c % A_STAGES;
      const uint32_t st = smem_u32(smem + (size_t)s * C::STAGE_BYTES);
      mbar_wait(&wfull[s], (c / S) & 1);
      if (DBG >= 3) {
        mbar_wait(&aempty[t], ((c / A_STAGES) & 1) ^ 1);
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&afull[t]);
          mbar_arrive(&sempty[s]);
        }
        continue;
      }

      uint4 v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t off = row_off + i * 16;
        v[i] = lds128(st + (off ^ (((off >> 7) & 7) << 4)));
      }
      uint32_t sz[4];  // (s,z) of the group each 32-k word belongs to
      {
        const uint32_t szb = st + W_BYTES + C::X_BYTES + r * 4;
#pragma unroll
        for (int wd = 0; wd < 4; ++wd) sz[wd] = lds32(szb + (wd >> gshift) * 512);
      }

      mbar_wait(&aempty[t], ((c / A_STAGES) & 1) ^ 1);
      tc_fence_after();

      uint32_t out[64];
#pragma unroll
      for (int wd = 0; wd < 4; ++wd) {
        const uint32_t szw = sz[wd];
        const uint32_t s_bits = __byte_perm(szw, szw, 0x1010);
        const uint32_t z_bits = __byte_perm(szw, szw, 0x3232);
        const __nv_bfloat162 s2 = *reinterpret_cast<const __nv_bfloat162*>(&s_bits);
        const __nv_bfloat162 z2 = *reinterpret_cast<const __nv_bfloat162*>(&z_bits);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t word = (wd == 0) ? v[i].x : (wd == 1) ? v[i].y : (wd == 2) ? v[i].z : v[i].w;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t m = ((word >> (4 * e)) & 0x000F000Fu) | 0x43004300u;
            out[16 * wd + i + 4 * e] = (DBG == 1) ? (word + e) : deq_pair(m, s2, z2);
          }
        }
      }
      const uint32_t a_t = lane_taddr + C::A_COL0 + t * A_COLS;
      if (DBG == 2) {
        uint32_t x = 0;
#pragma unroll
        for (int i = 0; i < 64; ++i) x ^= out[i];
        if (x == 0x12345678u) p.ws_sem[1000 + threadIdx.x] = x;  // keep the math alive
      } else {
        tmem_st_x32(a_t, out);
        tmem_st_x32(a_t + 32, out + 32);
        tc_wait_st();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&afull[t]);
        mbar_arrive(&sempty[s]);
      }
    }

    // ------------------------------------------------------------ epilogue
    pdl_wait();  // y / workspace may still be read by the previous kernel
    mbar_wait(dfull, 0);
    tc_fence_after();
    constexpr int HALF = N_MMA / 2;  // each warpgroup takes half of the token columns
    float acc[HALF];
#pragma unroll
    for (int j = 0; j < HALF; j += 8) {
      uint32_t rr[8];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
          : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]),
            "=r"(rr[6]), "=r"(rr[7])
          : "r"(lane_taddr + C::D_COL + wg * HALF + j)
          : "memory");
      tc_wait_ld();
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[j + q] = __uint_as_float(rr[q]);
    }
    const int n = n0 + r;
    const int mbase = m0 + wg * HALF;
    if (p.splits == 1) {
      if (n < p.N_out) {
        const float b = p.bias ? __bfloat162float(p.bias[n]) : 0.f;
#pragma unroll
        for (int j = 0; j < HALF; ++j)
          if (mbase + j < p.M)
            p.y[(size_t)(mbase + j) * p.N_out + n] = __float2bfloat16_rn(acc[j] + b);
      }
    } else {
      const int tile_lin = m_blk * gridDim.x + n_tile;
      float* part = p.ws_partial + ((size_t)tile_lin * p.splits + split) * (N_MMA * ROWS);
#pragma unroll
      for (int j = 0; j < HALF; ++j)
        if (mbase + j < p.M) __stcg(&part[(wg * HALF + j) * ROWS + r], acc[j]);
      __threadfence();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(&p.ws_sem[tile_lin], 1u);
        *flag_slot = (prev == (unsigned)p.splits - 1) ? 1u : 0u;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (*flag_slot) {
        __threadfence();
        const float* base = p.ws_partial + (size_t)tile_lin * p.splits * (N_MMA * ROWS);
        if (n < p.N_out) {
          const float b = p.bias ? __bfloat162float(p.bias[n]) : 0.f;
          for (int j = 0; j < HALF; ++j) {
            if (mbase + j >= p.M) break;
            float sum = 0.f;
            for (int sp = 0; sp < p.splits; ++sp)
              sum += __ldcg(&base[(size_t)sp * (N_MMA * ROWS) + (wg * HALF + j) * ROWS + r]);
            p.y[(size_t)(mbase + j) * p.N_out + n] = __float2bfloat16_rn(sum + b);
          }
        }
        if (threadIdx.x == 0) p.ws_sem[tile_lin] = 0;  // restore for the next launch
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------
// Reference-grade CUDA-core kernel (impl = 2): one warp per weight row, lanes stride over
// k-tiles, fp32 accumulation.  Slow; used as an on-device cross-check and for odd shapes.
template <int MT>
__global__ void int4_linear_simple_kernel(const __nv_bfloat16* __restrict__ x,
                                          const int32_t* __restrict__ qdata,
                                          const __nv_bfloat16* __restrict__ sz,
                                          const __nv_bfloat16* __restrict__ bias,
                                          __nv_bfloat16* __restrict__ y, int M, int N, int N_out,
                                          int K, int g) {
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
              if (m0 + m < M) acc[m] += __bfloat162float(x[(size_t)(m0 + m) * K + k]) * wf;
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

template <int N_MMA, int DBG = 0>
static int launch_tc(const uint16_t* x, int M, int K, const int32_t* qdata, const uint16_t* sz,
                     int g, int N, const uint16_t* bias, uint16_t* y, int N_out, void* ws,
                     size_t ws_bytes, cudaStream_t stream) {
  using C = Cfg<N_MMA>;
  const int KT = K / 128;
  CUtensorMap tm_w, tm_sz, tm_x;
  {
    const uint64_t dims[3] = {32, (uint64_t)4 * KT, (uint64_t)N / 8};
    const uint64_t str[2] = {128, (uint64_t)KT * 512};
    const uint32_t box[3] = {32, 4, 16};
    int rc = make_tmap(&tm_w, CU_TENSOR_MAP_DATA_TYPE_INT32, 3, qdata, dims, str, box,
                       CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  const int gpc = g <= 128 ? 128 / g : 1;
  {
    const uint64_t dims[2] = {(uint64_t)N, (uint64_t)K / g};
    const uint64_t str[1] = {(uint64_t)N * 4};
    const uint32_t box[2] = {128, (uint32_t)gpc};
    int rc = make_tmap(&tm_sz, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, sz, dims, str, box,
                       CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    const uint64_t str[1] = {(uint64_t)K * 2};
    const uint32_t box[2] = {64, (uint32_t)N_MMA};
    int rc = make_tmap(&tm_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x, dims, str, box,
                       CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  const int n_tiles = ceil_div(N_out, ROWS);
  const int m_blocks = ceil_div(M, N_MMA);
  int splits = sm_count() / (n_tiles * m_blocks);
  if (splits < 1) splits = 1;
  if (splits > MAX_SPLITS) splits = MAX_SPLITS;
  if (splits > KT) splits = KT;
  Params p;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.ws_sem = reinterpret_cast<unsigned int*>(ws);
  p.ws_partial = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + 64 * 1024);
  p.M = M; p.N = N; p.N_out = N_out; p.K = K; p.group_size = g; p.splits = splits;
  if (splits > 1) {
    const size_t need = 64 * 1024 + (size_t)n_tiles * m_blocks * splits * N_MMA * ROWS * 4;
    if (!ws || ws_bytes < need || (size_t)n_tiles * m_blocks * 4 > 64 * 1024)
      return fail(AO_ERR_WORKSPACE, "int4 linear: workspace too small (%zu < %zu)", ws_bytes, need);
  }
  auto kern = int4_linear_tc_kernel<N_MMA, DBG>;
  static bool attr_set = false;
  if (!attr_set) {
    AO_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)C::SMEM_BYTES));
    attr_set = true;
  }
  AO_CUDA_CHECK(launch(kern, dim3(n_tiles, splits, m_blocks), dim3(NUM_THREADS), C::SMEM_BYTES,
                       stream, pdl_enabled(), tm_w, tm_sz, tm_x, p));
  return AO_OK;
}

}  // namespace int4k
}  // namespace ao

extern "C" int ao_int4_tilepacked_linear(const uint16_t* x, int M, int K, const int32_t* qdata,
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
  if (M == 0) return AO_OK;
  AO_REQUIRE(x && qdata && scale_and_zero && y, "int4 linear: null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (impl == 2) {
    constexpr int MT = 8;
    dim3 grid(ceil_div(N, 8), ceil_div(M, MT));
    AO_CUDA_CHECK(launch(int4k::int4_linear_simple_kernel<MT>, grid, dim3(256), 0, st, false,
                         reinterpret_cast<const __nv_bfloat16*>(x), qdata,
                         reinterpret_cast<const __nv_bfloat16*>(scale_and_zero),
                         reinterpret_cast<const __nv_bfloat16*>(bias),
                         reinterpret_cast<__nv_bfloat16*>(y), M, N, N_out, K, group_size));
    return AO_OK;
  }
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("AO_B200_INT4_DBG");
    dbg = e ? atoi(e) : 0;
  }
  if (dbg > 0 && M <= 32) {
#define AO_DBG_CASE(NM, D)                                                                      \
  if (dbg == D)                                                                                 \
    return int4k::launch_tc<NM, D>(x, M, K, qdata, scale_and_zero, group_size, N, bias, y, N_out, \
                                   workspace, workspace_bytes, st);
    if (M <= 16) { AO_DBG_CASE(16, 1) AO_DBG_CASE(16, 2) AO_DBG_CASE(16, 3) AO_DBG_CASE(16, 4) AO_DBG_CASE(16, 5) AO_DBG_CASE(16, 6) }
    else { AO_DBG_CASE(32, 1) AO_DBG_CASE(32, 2) AO_DBG_CASE(32, 3) AO_DBG_CASE(32, 4) AO_DBG_CASE(32, 5) AO_DBG_CASE(32, 6) }
#undef AO_DBG_CASE
  }
  if (M <= 16)
    return int4k::launch_tc<16>(x, M, K, qdata, scale_and_zero, group_size, N, bias, y, N_out,
                                workspace, workspace_bytes, st);
  if (M <= 32)
    return int4k::launch_tc<32>(x, M, K, qdata, scale_and_zero, group_size, N, bias, y, N_out,
                                workspace, workspace_bytes, st);
  if (M <= 64)
    return int4k::launch_tc<64>(x, M, K, qdata, scale_and_zero, group_size, N, bias, y, N_out,
                                workspace, workspace_bytes, st);
  return int4k::launch_tc<128>(x, M, K, qdata, scale_and_zero, group_size, N, bias, y, N_out,
                               workspace, workspace_bytes, st);
}
