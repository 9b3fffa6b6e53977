// Swap-AB streaming GEMMs for 8-bit / 4-bit operands on tcgen05 (sm_100a):
//   int8 x int8 -> int32  (kind::i8)         replaces aten._int_mm + the scale epilogue
//   e4m3 x e4m3 -> f32    (kind::f8f6f4)     replaces torch._scaled_mm rowwise
//   mxfp8 block-32 e8m0   (kind::mxf8f6f4)   replaces torch._scaled_mm block-scaled
//   nvfp4 block-16 e4m3   (kind::mxf4nvf4)   replaces torch._scaled_mm fp4 + pts/bias kernels
// Reference call sites: int8/kernels.py:18-76,114-144 + int8_tensor.py:305-359;
// float8/inference.py:86-123; mx_formats/mx_tensor.py:759-843; nvfp4_tensor.py:487-578.
//
// Decode-shaped (M <= 128 tokens per block).  The weight matrix W[N,K] (K-major, exactly the
// stored qdata) is the UMMA A operand: 128 output features per CTA; the activations are the B
// operand (N_MMA tokens).  TMA (SWIZZLE_128B) streams 128-byte-wide K blocks of W through a
// deep smem ring straight into tcgen05.mma (SS mode) -- no register staging at all; block
// scales go smem -> TMEM with tcgen05.cp.  Epilogue: TMEM -> registers -> scale/bias -> bf16.
// Persistent stream-K (streamk.cuh): (tile, K-chunk) units split evenly over the CTAs; tiles shared by several
// CTAs are finished by their owner CTA from the contributors' published 32-bit partials in CTA order
// (deterministic; int32 for int8: stays exact).
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"
#include "streamk.cuh"

namespace ao {
namespace lowp {

enum Kind { KIND_I8 = 0, KIND_F8 = 1, KIND_MXF8 = 2, KIND_NVF4 = 3 };

using streamk::ROWS;
constexpr int KB = 128;             // bytes of K per stage per row (128 elems for 8-bit, 256 for fp4)
constexpr int A_BYTES = ROWS * KB;  // 16 KiB
constexpr int EPI_WARP0 = 0, TMA_WARP = 4, MMA_WARP = 5;
constexpr int NUM_THREADS = 192;

template <int KIND, int N_MMA>
struct Cfg {
  static constexpr bool BLOCK_SCALED = (KIND == KIND_MXF8 || KIND == KIND_NVF4);
  static constexpr int B_BYTES = N_MMA * KB;
  // scale tiles per stage: mxfp8 1 blocked tile (128 rows x 4 sf), nvfp4 4 tiles (16 sf / row)
  static constexpr int SF_TILES = KIND == KIND_MXF8 ? 1 : (KIND == KIND_NVF4 ? 4 : 0);
  static constexpr int SF_BYTES = SF_TILES * 512;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES + 2 * SF_BYTES + (BLOCK_SCALED ? (1024 - (2 * SF_BYTES) % 1024) % 1024 : 0);
  static constexpr int BUDGET = N_MMA <= 64 ? 104 * 1024 : 168 * 1024;
  static constexpr int STAGES = BUDGET / STAGE_BYTES;
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 + 1024;
  // TMEM: two accumulators, then the scale factors (double-buffered by chunk parity)
  static constexpr int D_COL0 = 0, D_COL1 = N_MMA;
  static constexpr int SF_COLS = SF_TILES * 4;  // columns per operand per chunk
  static constexpr int SFA_COL = N_MMA <= 64 ? 128 : 256;
  static constexpr int SFB_COL = SFA_COL + 32;
  static constexpr int TMEM_COLS = N_MMA <= 64 ? (BLOCK_SCALED ? 256 : (N_MMA <= 16 ? 32 : (N_MMA <= 32 ? 64 : 128))) : (BLOCK_SCALED ? 512 : 256);
  static constexpr int MMA_PER_STAGE = 4;  // K bytes per MMA = 32 (8-bit: 32 elems, fp4: 64)
};

struct Params {
  const float* x_scale;   // [M]     (rowwise kinds) or a_pts scalar (nvfp4) or null
  const float* w_scale;   // [N]     (rowwise kinds) or b_pts scalar (nvfp4) or null
  const __nv_bfloat16* bias;
  __nv_bfloat16* y;       // bf16 out [M, N]  (null when i32_out is set)
  int32_t* i32_out;       // raw int32 accumulators [M, N] (ao_int8_mm_i32)
  float* ws_partial;      // [grid][N_MMA*128] 32-bit partials (int32 bits for int8): CTA b's CONTRIB partial
  unsigned int* ws_flag;  // [grid] CTA b's partial is published (streamk.cuh)
  int M, N, K;            // K in ELEMENTS
  int n_tiles, m_blocks, KT;  // KT = chunks of 128 K-bytes
  int sf_col_blocks_w;    // number of 4-wide scale column blocks per row block (blocked layout)
  int sf_col_blocks_x;
};

using streamk::cta_of_unit;
using streamk::unit_begin;

// Persistent stream-K kernel (same work split / fix-up protocol as ts_gemm.cuh, SS-mode MMAs):
// warps 0-3 epilogue, warp 4 TMA producer, warp 5 MMA issuer.
template <int KIND, int N_MMA>
__global__ void __launch_bounds__(NUM_THREADS, (N_MMA <= 64 ? 2 : 1))
lowp_linear_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_x,
                   const uint8_t* __restrict__ w_sf, const uint8_t* __restrict__ x_sf,
                   const Params p) {
  using C = Cfg<KIND, N_MMA>;
  constexpr int S = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * C::STAGE_BYTES);
  uint64_t* wfull = bars;            // [S] weights (+ weight scales)
  uint64_t* xfull = wfull + S;       // [S] activations (+ activation scales)
  uint64_t* sempty = xfull + S;      // [S] MMA commit
  uint64_t* dfull = sempty + S;      // [2]
  uint64_t* dempty = dfull + 2;      // [2] 4 epilogue warps
  uint64_t* fok = dempty + 2;        // [1] the contributors' flags of an OWNER last segment have been seen (TMA warp polls)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(fok + 1);

  // warp index through a shuffle => known warp-uniform: the single-thread roles are warp-uniform loops with only
  // the TMA / tcgen05 instructions under elect.sync, so their operands live in uniform registers (with the whole
  // loop under `lane == 0` ptxas wraps every UTCHMMA / UTMALDG in an elect-broadcast loop, ~2.5x slower issue)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int G = gridDim.x, b = blockIdx.x;
  const long long U = (long long)p.n_tiles * p.m_blocks * p.KT;
  const int u0 = unit_begin(b, U, G), u1 = unit_begin(b + 1, U, G);
  const int nunits = u1 - u0;
  const streamk::Walk walk(u0, nunits, p.KT);
  auto tile_of = [&](int i) { return (u0 + i) / p.KT; };
  auto kc_of = [&](int i) { return (u0 + i) % p.KT; };

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&wfull[i], 1);
      mbar_init(&xfull[i], 1);
      mbar_init(&sempty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&dfull[i], 1);
      mbar_init(&dempty[i], 4);
    }
    mbar_init(fok, 1);
    fence_barrier_init();
  }
  if (warp == TMA_WARP && lane == 0) {
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_x);
  }
  if (warp == MMA_WARP) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_launch_dependents();

  if (warp == TMA_WARP) {
    {
      const uint64_t pol_w = policy_evict_first();
      const uint64_t pol_x = policy_evict_last();
      auto issue_w = [&](int i) {
        const int s = i % S, n_tile = tile_of(i) % p.n_tiles, kc = kc_of(i);
        uint8_t* st = smem + (size_t)s * C::STAGE_BYTES;
        mbar_expect_tx(&wfull[s], A_BYTES + C::SF_BYTES);
        tma_load_2d(st, &tm_w, &wfull[s], kc * KB, n_tile * ROWS, pol_w);
        if (C::BLOCK_SCALED) {
          // blocked scale tiles of this (row block, k chunk): SF_TILES consecutive 512-byte tiles
          const uint8_t* src = w_sf + ((size_t)n_tile * p.sf_col_blocks_w + (size_t)kc * C::SF_TILES) * 512;
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                  smem_u32(st + A_BYTES + C::B_BYTES)),
              "l"(src), "r"(C::SF_BYTES), "r"(smem_u32(&wfull[s]))
              : "memory");
        }
      };
      auto issue_x = [&](int i) {
        const int s = i % S, m0 = (tile_of(i) / p.n_tiles) * N_MMA, kc = kc_of(i);
        uint8_t* st = smem + (size_t)s * C::STAGE_BYTES;
        mbar_expect_tx(&xfull[s], C::B_BYTES + C::SF_BYTES);
        tma_load_2d(st + A_BYTES, &tm_x, &xfull[s], kc * KB, m0, pol_x);
        if (C::BLOCK_SCALED) {
          const uint8_t* src = x_sf + ((size_t)(m0 / 128) * p.sf_col_blocks_x + (size_t)kc * C::SF_TILES) * 512;
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                  smem_u32(st + A_BYTES + C::B_BYTES + C::SF_BYTES)),
              "l"(src), "r"(C::SF_BYTES), "r"(smem_u32(&xfull[s]))
              : "memory");
        }
      };
      const int pre = nunits < S ? nunits : S;
      for (int i = 0; i < pre; ++i) {
        if (elect_one()) issue_w(i);   // weights never depend on the previous kernel
        __syncwarp();
      }
      pdl_wait();
      for (int i = 0; i < pre; ++i) {
        if (elect_one()) issue_x(i);
        __syncwarp();
      }
      for (int i = S; i < nunits; ++i) {
        mbar_wait(&sempty[i % S], ((i / S) & 1) ^ 1);
        if (elect_one()) {
          issue_w(i);
          issue_x(i);
        }
        __syncwarp();
      }
      if (walk.seg_kind(walk.nseg - 1) == streamk::SEG_OWNER) {
        // this warp is done a ring depth before the CTA is: it polls the contributors' flags (ld.acquire.gpu) and passes
        // the result on through a cta-scope barrier, so that L2 round trip runs under the last MMAs instead of between
        // them and the owner's gather (same scheme as ts_gemm.cuh)
        const int tile_l = walk.seg_tile(walk.nseg - 1);
        const int n_oth = cta_of_unit((long long)tile_l * p.KT + p.KT - 1, U, G) - b;
        for (int base = 0; base < n_oth; base += 32)
          if (base + lane < n_oth)
            while (streamk::ld_acquire_u32(p.ws_flag + b + 1 + base + lane) == 0u) __nanosleep(64);
        __syncwarp();
        if (elect_one()) mbar_arrive(fok);
        __syncwarp();
      }
    }
  } else if (warp == MMA_WARP) {
    constexpr uint32_t idesc =
        KIND == KIND_I8   ? make_idesc(2 /*s32*/, 1 /*int8*/, 1 /*int8*/, ROWS, N_MMA)
        : KIND == KIND_F8 ? make_idesc(1 /*f32*/, 0 /*e4m3*/, 0 /*e4m3*/, ROWS, N_MMA)
        : KIND == KIND_MXF8 ? make_idesc_bs(0 /*e4m3*/, 0 /*e4m3*/, 1 /*ue8m0*/, ROWS, N_MMA)
                            : make_idesc_bs(1 /*e2m1*/, 1 /*e2m1*/, 0 /*ue4m3*/, ROWS, N_MMA);
    int seg = 0;
    for (int i = 0; i < nunits; ++i) {
      const int s = i % S;
      const bool first = (i == 0) || (kc_of(i) == 0);
      const bool last = (i == nunits - 1) || (kc_of(i) == p.KT - 1);
      const int buf = seg & 1;
      if (first) {
        mbar_wait(&dempty[buf], ((seg >> 1) & 1) ^ 1);
        tc_fence_after();
      }
      mbar_wait(&wfull[s], (i / S) & 1);
      mbar_wait(&xfull[s], (i / S) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t ab = smem_u32(smem + (size_t)s * C::STAGE_BYTES);
        const uint32_t bb = ab + A_BYTES;
        const uint32_t d_t = tmem_base + (buf ? C::D_COL1 : C::D_COL0);
        uint32_t sfa_t = 0, sfb_t = 0;
        if (C::BLOCK_SCALED) {
          const int sb = i & 1;
          sfa_t = tmem_base + C::SFA_COL + sb * C::SF_COLS;
          sfb_t = tmem_base + C::SFB_COL + sb * C::SF_COLS;
#pragma unroll
          for (int tI = 0; tI < C::SF_TILES; ++tI) {
            // 512-byte tile = 32 rows x 16 bytes, contiguous: one 8x16B core matrix per 128 B
            const uint64_t da = umma_desc_k_noswz(bb + C::B_BYTES + tI * 512, 16, 128);
            const uint64_t db = umma_desc_k_noswz(bb + C::B_BYTES + C::SF_BYTES + tI * 512, 16, 128);
            tc_cp_32x128b_warpx4(sfa_t + tI * 4, da);
            tc_cp_32x128b_warpx4(sfb_t + tI * 4, db);
          }
        }
#pragma unroll
        for (int kk = 0; kk < C::MMA_PER_STAGE; ++kk) {
          const uint64_t adesc = umma_desc_k_sw128(ab + kk * 32);
          const uint64_t bdesc = umma_desc_k_sw128(bb + kk * 32);
          const uint32_t acc = (!first || kk > 0) ? 1u : 0u;
          if (KIND == KIND_I8) mma_ss_i8(d_t, adesc, bdesc, idesc, acc);
          else if (KIND == KIND_F8) mma_ss_f8f6f4(d_t, adesc, bdesc, idesc, acc);
          else if (KIND == KIND_MXF8)
            // one scale byte per row per MMA: byte kk of the tile's column (sf_id fields)
            mma_ss_mxf8f6f4(d_t, adesc, bdesc, idesc | ((uint32_t)kk << 29) | ((uint32_t)kk << 4), acc, sfa_t, sfb_t);
          else
            // four scale bytes per row per MMA: tile kk
            mma_ss_mxf4nvf4_b16(d_t, adesc, bdesc, idesc, acc, sfa_t + kk * 4, sfb_t + kk * 4);
        }
        tc_commit(&sempty[s]);
        if (last) tc_commit(&dfull[buf]);
      }
      __syncwarp();
      if (last) ++seg;
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 0..3)
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    pdl_wait();
    for (int seg = 0; seg < walk.nseg; ++seg) {
      const int tile = walk.seg_tile(seg);
      const int kind = walk.seg_kind(seg);
      const int buf = seg & 1;
      mbar_wait(&dfull[buf], (seg >> 1) & 1);
      tc_fence_after();
      const int n_tile = tile % p.n_tiles, m_blk = tile / p.n_tiles;
      const int n = n_tile * ROWS + r, m0 = m_blk * N_MMA;
      const uint32_t d_t = lane_taddr + (buf ? C::D_COL1 : C::D_COL0);
      float sw = 1.f, bias = 0.f;
      if (n < p.N) {
        if (KIND == KIND_I8 || KIND == KIND_F8) sw = p.w_scale ? p.w_scale[n] : 1.f;
        if (KIND == KIND_NVF4) sw = (p.x_scale ? *p.x_scale : 1.f) * (p.w_scale ? *p.w_scale : 1.f);
        if (p.bias) bias = __bfloat162float(p.bias[n]);
      }
      // finalise one accumulator value (raw 32 bits: s32 for int8, f32 otherwise) into the output
      auto emit = [&](int m, uint32_t raw_i, float raw_f) {
        if (KIND == KIND_I8) {
          if (p.i32_out) {
            p.i32_out[(size_t)m * p.N + n] = (int32_t)raw_i;
            return;
          }
          // int8/kernels.py:143-144 + int8_tensor.py:315-359: bf16 round between the scales
          const float t = __bfloat162float(__float2bfloat16_rn((float)(int32_t)raw_i * p.x_scale[m]));
          p.y[(size_t)m * p.N + n] = __float2bfloat16_rn(t * sw + bias);
        } else if (KIND == KIND_F8) {
          p.y[(size_t)m * p.N + n] = __float2bfloat16_rn(raw_f * (p.x_scale[m] * sw) + bias);
        } else if (KIND == KIND_MXF8) {
          p.y[(size_t)m * p.N + n] = __float2bfloat16_rn(raw_f + bias);
        } else {
          p.y[(size_t)m * p.N + n] = __float2bfloat16_rn(raw_f * sw + bias);
        }
      };
      if (kind == streamk::SEG_FULL) {
#pragma unroll
        for (int j = 0; j < N_MMA; j += 16) {
          uint32_t rr[16];
          tmem_ld_x16(d_t + j, rr);
          tc_wait_ld();
          if (n < p.N) {
#pragma unroll
            for (int q = 0; q < 16; ++q)
              if (m0 + j + q < p.M) emit(m0 + j + q, rr[q], __uint_as_float(rr[q]));
          }
        }
      } else if (kind == streamk::SEG_CONTRIB) {
        // publish the partial (column-major slot: coalesced across the 128 rows), then raise this CTA's flag
        uint32_t* slot = reinterpret_cast<uint32_t*>(p.ws_partial) + (size_t)b * (N_MMA * ROWS) + r;
#pragma unroll
        for (int j = 0; j < N_MMA; j += 16) {
          uint32_t rr[16];
          tmem_ld_x16(d_t + j, rr);
          tc_wait_ld();
#pragma unroll
          for (int q = 0; q < 16; ++q)
            if (m0 + j + q < p.M) __stcg(&slot[(j + q) * ROWS], rr[q]);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");   // all 128 rows stored (cta-scope order) ...
        if (threadIdx.x == EPI_WARP0 * 32) streamk::st_release_u32(p.ws_flag + b, 1u);   // ... one gpu-scope release
      } else {
        // OWNER: own partial (TMEM) + the partials of CTAs b+1 .. b_last in that order (fixed: deterministic)
        const int b_last = cta_of_unit((long long)tile * p.KT + p.KT - 1, U, G);
        const int n_oth = b_last - b;
        const uint32_t* slot0 = reinterpret_cast<const uint32_t*>(p.ws_partial) + (size_t)(b + 1) * (N_MMA * ROWS) + r;
        mbar_wait(fok, 0);   // flags seen by the TMA warp (above)
#pragma unroll 1
        for (int j0 = 0; j0 < N_MMA; j0 += 16) {
          if (m0 + j0 >= p.M) break;
          uint32_t own[16];
          tmem_ld_x16(d_t + j0, own);
          tc_wait_ld();
          float vf[16];
          int32_t vi[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            vf[q] = __uint_as_float(own[q]);
            vi[q] = (int32_t)own[q];
          }
#pragma unroll 1
          for (int c0 = 0; c0 < n_oth; c0 += 2) {
            // two contributors x 16 columns of loads in flight (the gather is a chain of L2 round trips)
            uint32_t t0[16], t1[16];
            const uint32_t* s0 = slot0 + (size_t)c0 * (N_MMA * ROWS) + (size_t)j0 * ROWS;
            const uint32_t* s1 = s0 + (size_t)(N_MMA * ROWS);
            const bool has1 = c0 + 1 < n_oth;
#pragma unroll
            for (int q = 0; q < 16; ++q) t0[q] = (n < p.N && m0 + j0 + q < p.M) ? __ldcg(s0 + q * ROWS) : 0u;
#pragma unroll
            for (int q = 0; q < 16; ++q) t1[q] = (has1 && n < p.N && m0 + j0 + q < p.M) ? __ldcg(s1 + q * ROWS) : 0u;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              vi[q] += (int32_t)t0[q] + (int32_t)t1[q];
              vf[q] = (vf[q] + __uint_as_float(t0[q])) + __uint_as_float(t1[q]);
            }
          }
          if (n < p.N) {
#pragma unroll
            for (int q = 0; q < 16; ++q)
              if (m0 + j0 + q < p.M) emit(m0 + j0 + q, (uint32_t)vi[q], vf[q]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dempty[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == EPI_WARP0 && walk.seg_kind(walk.nseg - 1) == streamk::SEG_OWNER) {
    // the owner has consumed its contributors' partials: re-arm their flags for the next launch
    const int b_last = cta_of_unit((long long)walk.seg_tile(walk.nseg - 1) * p.KT + p.KT - 1, U, G);
    for (int c = b + 1 + lane; c <= b_last; c += 32) p.ws_flag[c] = 0u;
  }
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

template <int KIND, int N_MMA>
static int launch(const uint8_t* xq, const uint8_t* x_sf, const float* x_scale, int M, int K,
                  const uint8_t* wq, const uint8_t* w_sf, const float* w_scale, int N,
                  const uint16_t* bias, uint16_t* y, int32_t* i32_out, void* ws, size_t ws_bytes,
                  cudaStream_t stream) {
  using C = Cfg<KIND, N_MMA>;
  const int epb = KIND == KIND_NVF4 ? 2 : 1;
  const int k_bytes = K / epb;
  CUtensorMap tm_w, tm_x;
  {
    const uint64_t dims[2] = {(uint64_t)k_bytes, (uint64_t)N};
    const uint64_t str[1] = {(uint64_t)k_bytes};
    const uint32_t box[2] = {KB, ROWS};
    int rc = make_tmap(&tm_w, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, wq, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)k_bytes, (uint64_t)M};
    const uint64_t str[1] = {(uint64_t)k_bytes};
    const uint32_t box[2] = {KB, (uint32_t)N_MMA};
    int rc = make_tmap(&tm_x, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, xq, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  Params p{};
  p.x_scale = x_scale; p.w_scale = w_scale;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.i32_out = i32_out;
  p.ws_flag = reinterpret_cast<unsigned int*>(ws);
  p.ws_partial = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + streamk::WS_PARTIAL_OFF);
  p.M = M; p.N = N; p.K = K;
  p.n_tiles = ceil_div(N, ROWS);
  p.m_blocks = ceil_div(M, N_MMA);
  p.KT = ceil_div(k_bytes, KB);   // a K tail is zero-filled by TMA (out-of-bounds box elements) on both operands
  const int sf_per_row = KIND == KIND_MXF8 ? K / 32 : (KIND == KIND_NVF4 ? K / 16 : 0);
  p.sf_col_blocks_w = ceil_div(sf_per_row, 4);
  p.sf_col_blocks_x = ceil_div(sf_per_row, 4);
  const long long units = (long long)p.n_tiles * p.m_blocks * p.KT;
  // never fewer than 8 chunks per CTA: finer splits lengthen the split-tile reduction more than they shorten the
  // streaming phase (fp8, q|k|v projection: 12.6 -> 10.5 us at 1 token, 15.2 -> 13.5 us at 32 with 8 instead of 4; 16 is
  // worse again: profiles/r02_call_p.log)
  // two CTAs per SM where they fit (N_MMA <= 64): this kernel has no dequant phase, it is bounded by the bytes in
  // flight per SM (stages x 16-20 KB against the DRAM round trip), which a second resident CTA doubles
  const int per_sm = (N_MMA <= 64) ? (ts_ctas_per_sm() ? ts_ctas_per_sm() : 2) : 1;
  int grid = sm_count() * per_sm;
  const int min_units = ts_min_units() ? ts_min_units() : 8;
  if (units / min_units < grid) grid = units / min_units > 0 ? (int)(units / min_units) : 1;
  const size_t need = streamk::WS_PARTIAL_OFF + (size_t)grid * N_MMA * ROWS * 4;
  if (!ws || ws_bytes < need || (size_t)grid * 4 > streamk::WS_FLAGS_BYTES)
    return fail(AO_ERR_WORKSPACE, "lowp linear: workspace too small (%zu < %zu)", ws_bytes, need);
  auto kern = lowp_linear_kernel<KIND, N_MMA>;
  AO_CUDA_CHECK(ensure_dynamic_smem(reinterpret_cast<const void*>(kern), C::SMEM_BYTES));
  AO_CUDA_CHECK(ao::launch(kern, dim3(grid), dim3(NUM_THREADS), C::SMEM_BYTES, stream, pdl_enabled(), tm_w, tm_x,
                           w_sf, x_sf, p));
  return AO_OK;
}

template <int KIND>
static int dispatch(const uint8_t* xq, const uint8_t* x_sf, const float* x_scale, int M, int K,
                    const uint8_t* wq, const uint8_t* w_sf, const float* w_scale, int N,
                    const uint16_t* bias, uint16_t* y, int32_t* i32_out, void* ws, size_t ws_bytes,
                    void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (M <= 16 && KIND != KIND_MXF8 && KIND != KIND_NVF4)
    return launch<KIND, 16>(xq, x_sf, x_scale, M, K, wq, w_sf, w_scale, N, bias, y, i32_out, ws, ws_bytes, st);
  if (M <= 32)
    return launch<KIND, 32>(xq, x_sf, x_scale, M, K, wq, w_sf, w_scale, N, bias, y, i32_out, ws, ws_bytes, st);
  if (M <= 64)
    return launch<KIND, 64>(xq, x_sf, x_scale, M, K, wq, w_sf, w_scale, N, bias, y, i32_out, ws, ws_bytes, st);
  return launch<KIND, 128>(xq, x_sf, x_scale, M, K, wq, w_sf, w_scale, N, bias, y, i32_out, ws, ws_bytes, st);
}

}  // namespace lowp
}  // namespace ao

using namespace ao;

extern "C" int ao_int8_dyn_linear(const int8_t* xq, const float* x_scale, int M, int K,
                                  const int8_t* wq, const float* w_scale, int N,
                                  const uint16_t* bias, uint16_t* y, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  AO_REQUIRE(M >= 0 && K > 0 && N > 0, "int8 linear: bad sizes M=%d K=%d N=%d", M, K, N);
  AO_REQUIRE(K % 16 == 0, "int8 linear: K=%d must be a multiple of 16 (TMA row pitch)", K);
  AO_REQUIRE(N % 8 == 0, "int8 linear: N=%d must be a multiple of 8 (reference: int8/kernels.py:48-58)", N);
  if (M == 0) return AO_OK;
  AO_REQUIRE(xq && x_scale && wq && w_scale && y, "int8 linear: null pointer");
  return lowp::dispatch<lowp::KIND_I8>(reinterpret_cast<const uint8_t*>(xq), nullptr, x_scale, M, K,
                                       reinterpret_cast<const uint8_t*>(wq), nullptr, w_scale, N, bias,
                                       y, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int ao_int8_mm_i32(const int8_t* xq, int M, int K, const int8_t* wq, int N, int32_t* acc,
                              void* workspace, size_t workspace_bytes, void* stream) {
  AO_REQUIRE(M >= 0 && K > 0 && N > 0 && K % 16 == 0, "int8 mm: bad sizes M=%d K=%d N=%d (K%%16==0)", M, K, N);
  if (M == 0) return AO_OK;
  AO_REQUIRE(xq && wq && acc, "int8 mm: null pointer");
  return lowp::dispatch<lowp::KIND_I8>(reinterpret_cast<const uint8_t*>(xq), nullptr, nullptr, M, K,
                                       reinterpret_cast<const uint8_t*>(wq), nullptr, nullptr, N, nullptr,
                                       nullptr, acc, workspace, workspace_bytes, stream);
}

extern "C" int ao_fp8_rowwise_linear(const uint8_t* xq, const float* x_scale, int M, int K,
                                     const uint8_t* wq, const float* w_scale, int N,
                                     const uint16_t* bias, uint16_t* y, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  AO_REQUIRE(M >= 0 && K > 0 && N > 0, "fp8 linear: bad sizes M=%d K=%d N=%d", M, K, N);
  AO_REQUIRE(K % 16 == 0, "fp8 linear: K=%d must be a multiple of 16 (reference: quantization/utils.py:663-687)", K);
  AO_REQUIRE(N % 16 == 0, "fp8 linear: N=%d must be a multiple of 16 (reference: quantization/utils.py:663-687)", N);
  if (M == 0) return AO_OK;
  AO_REQUIRE(xq && x_scale && wq && w_scale && y, "fp8 linear: null pointer");
  return lowp::dispatch<lowp::KIND_F8>(xq, nullptr, x_scale, M, K, wq, nullptr, w_scale, N, bias, y, nullptr,
                                       workspace, workspace_bytes, stream);
}

extern "C" int ao_mxfp8_linear(const uint8_t* xq, const uint8_t* x_scale_blocked, int M, int K,
                               const uint8_t* wq, const uint8_t* w_scale_blocked, int N,
                               const uint16_t* bias, uint16_t* y, void* workspace,
                               size_t workspace_bytes, void* stream) {
  AO_REQUIRE(M >= 0 && K > 0 && N > 0, "mxfp8 linear: bad sizes M=%d K=%d N=%d", M, K, N);
  AO_REQUIRE(K % 32 == 0, "mxfp8 linear: K=%d must be a multiple of 32 (mx_tensor.py:244-246)", K);
  if (M == 0) return AO_OK;
  AO_REQUIRE(xq && x_scale_blocked && wq && w_scale_blocked && y, "mxfp8 linear: null pointer");
  return lowp::dispatch<lowp::KIND_MXF8>(xq, x_scale_blocked, nullptr, M, K, wq, w_scale_blocked, nullptr, N,
                                         bias, y, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int ao_nvfp4_linear(const uint8_t* xq, const uint8_t* x_scale_blocked, const float* a_pts,
                               int M, int K, const uint8_t* wq, const uint8_t* w_scale_blocked,
                               const float* b_pts, int N, const uint16_t* bias, uint16_t* y,
                               void* workspace, size_t workspace_bytes, void* stream) {
  AO_REQUIRE(M >= 0 && K > 0 && N > 0, "nvfp4 linear: bad sizes M=%d K=%d N=%d", M, K, N);
  AO_REQUIRE(K % 256 == 0, "nvfp4 linear: K=%d must be a multiple of 256", K);
  AO_REQUIRE(N % 16 == 0, "nvfp4 linear: N=%d must be a multiple of 16 (inference_workflow.py:248-251)", N);
  if (M == 0) return AO_OK;
  AO_REQUIRE(xq && x_scale_blocked && wq && w_scale_blocked && y, "nvfp4 linear: null pointer");
  return lowp::dispatch<lowp::KIND_NVF4>(xq, x_scale_blocked, a_pts, M, K, wq, w_scale_blocked, b_pts, N, bias,
                                         y, nullptr, workspace, workspace_bytes, stream);
}
