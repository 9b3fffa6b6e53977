// Persistent "dequantise-into-TMEM" decode GEMM for sub-byte weights with bf16 activations (sm_100a).
//
//   Y[M,N] = X[M,K] * W^[N,K]^T (+bias),   W^ produced in-kernel from 4-bit weights + group scales
//
// Used by int4_linear.cu (tinygemm tile-packed int4, W^ = bf16((q-8)s+z)) and
// nvfp4_weight_linear.cu (e2m1 * e4m3 block scale).  Structure:
//   * swap-AB: 128 weight rows = UMMA M, tokens = UMMA N (16..128), fp32 accumulator in TMEM
//   * persistent, one CTA per SM; the (n-tile, 128-k chunk) units of the whole GEMM are split EVENLY
//     over the CTAs ("stream-K"), so every SM issues the same number of MMAs whatever N/K are
//   * 16 warps in four warpgroups with register redistribution (setmaxnreg):
//       WG0, WG1 (warps 0-7, 96 regs): dequant, alternating chunks: ld.shared (conflict-free through the
//                  TMA swizzle) -> unpack/scale in bf16x2 -> tcgen05.st of the bf16 A operand into TMEM
//       WG2 (warps 8-11, 32 regs)    : epilogue: tcgen05.ld of a finished accumulator (double-buffered,
//                  overlaps the next tile's MMAs), split-tile fix-up through an fp32 workspace, bias, store
//       WG3 (warp 12 TMA producer, warp 13 MMA issuer, warps 14-15 idle; 32 regs)
//     64 registers/thread at launch + 256 TMEM columns + ~105 KB smem => two CTAs fit on an SM, so the next
//     linear's CTA (PDL) is already resident and prefetching weights while this one drains
//   * tiles split across CTAs are reduced deterministically: every CTA writes its partial, bumps the
//     tile's unit counter, and whoever completes the count sums the partials in CTA order
//   * PDL: griddepcontrol.launch_dependents at start; weights are prefetched before
//     griddepcontrol.wait, only activations / outputs / workspace wait for the previous kernel.
#pragma once
#include <cuda_bf16.h>

#include "common.h"
#include "ptx.cuh"

namespace ao {
namespace tsg {

constexpr int ROWS = 128;
constexpr int KCHUNK = 128;
constexpr int W_BYTES = ROWS * KCHUNK / 2;  // 8 KiB of 4-bit weights per chunk
constexpr int AUX_BYTES = 2048;             // scales per chunk (<= 2 KiB), 1 KiB aligned slot
constexpr int A_COLS = 64;                  // TMEM columns of one bf16 A stage (128 k / 2)
constexpr int DEQ_WARPS = 8, EPI_WARP0 = 8, TMA_WARP = 12, MMA_WARP = 13;
constexpr int NUM_THREADS = 16 * 32;
constexpr int REGS_DEQ = 96, REGS_OTHER = 32;

template <int N_MMA>
struct Cfg {
  static constexpr int X_BYTES = 2 * N_MMA * 128;
  static constexpr int STAGE_BYTES = W_BYTES + X_BYTES + AUX_BYTES;
  static constexpr int BUDGET = N_MMA <= 64 ? 104 * 1024 : 172 * 1024;
  static constexpr int STAGES = BUDGET / STAGE_BYTES;
  static constexpr int TMEM_COLS = N_MMA <= 64 ? 256 : 512;
  static constexpr int D_COL0 = 0, D_COL1 = N_MMA;
  static constexpr int A_COL0 = N_MMA <= 32 ? 64 : 2 * N_MMA;
  static constexpr int A_STAGES = (TMEM_COLS - A_COL0) / A_COLS;  // 3, 2 or 4
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 + 1024;
};

struct Params {
  const __nv_bfloat16* bias;
  const float* row_scale;   // optional per-token scale [M] applied in the epilogue (fp8 activations)
  const float* out_scale;   // optional device scalar applied in the epilogue (nvfp4 per-tensor scale)
  __nv_bfloat16* y;         // [M, N_out]
  float* ws_partial;        // [grid][2][N_MMA*128]
  unsigned int* ws_sem;     // [tiles]
  const uint8_t* aux_base;  // format-specific scale base pointer (nvfp4: blocked scales)
  int M, N, N_out, K, group_size;
  int n_tiles, m_blocks, KT;   // KT = K/128
  int aux_col_blocks;
  int flags;  // bring-up switches (AO_B200_TS_FLAGS): 1 = dequant warpgroups on the high warp ids, 2 = sleep in TMA/MMA waits
  unsigned long long* timeline;  // debug: per-CTA [8] timestamps (AO_B200_TIMELINE=1), else null
};

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}

template <int R>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }

// the eight K=16 MMAs of one 128-k chunk in one straight-line block: A stage at TMEM column a0 (8 columns per
// MMA), B = two 64-k swizzle atoms (descriptors b_lo / b_hi, +32 B = +2 in the address field per MMA)
__device__ __forceinline__ void mma_chunk_ts_f16(uint32_t d, uint32_t a0, uint64_t b_lo, uint64_t b_hi,
                                                 uint32_t idesc, uint32_t accumulate_first) {
  asm volatile(
      "{\n\t"
      ".reg .pred p0, pt;\n\t"
      ".reg .b32 a1, a2, a3, a4, a5, a6, a7;\n\t"
      ".reg .b64 b1, b2, b3, b5, b6, b7;\n\t"
      "setp.ne.b32 p0, %5, 0;\n\t"
      "setp.eq.b32 pt, 0, 0;\n\t"
      "add.u32 a1, %1, 8;  add.u32 a2, %1, 16; add.u32 a3, %1, 24; add.u32 a4, %1, 32;\n\t"
      "add.u32 a5, %1, 40; add.u32 a6, %1, 48; add.u32 a7, %1, 56;\n\t"
      "add.u64 b1, %2, 2; add.u64 b2, %2, 4; add.u64 b3, %2, 6;\n\t"
      "add.u64 b5, %3, 2; add.u64 b6, %3, 4; add.u64 b7, %3, 6;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %4, p0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], b1, %4, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], b2, %4, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], b3, %4, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [a4], %3, %4, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [a5], b5, %4, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [a6], b6, %4, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [a7], b7, %4, pt;\n\t"
      "}\n" ::"r"(d),
      "r"(a0), "l"(b_lo), "l"(b_hi), "r"(idesc), "r"(accumulate_first)
      : "memory");
}

// unit range of CTA b: [U*b/G, U*(b+1)/G)
__device__ __forceinline__ int unit_begin(int b, long long U, int G) { return (int)((U * b) / G); }
__device__ __forceinline__ int cta_of_unit(int u, long long U, int G) {
  return (int)((((long long)(u + 1)) * G + U - 1) / U) - 1;
}

// Fmt policy:
//   static void issue_w(tm_w, tm_aux, p, w smem dst, aux smem dst, full barrier, n_tile, kc, policy)  (one thread)
//   static uint32_t w_tx_bytes(p)
//   static void dequant(p, w smem, aux smem, row r, out[64])      (128 threads; out[c] = bf16x2 of k = 2c, 2c+1)
// TL = true compiles the per-CTA phase-timestamp instrumentation in (bring-up builds only).
template <class Fmt, int N_MMA, bool TL = false>
__global__ void __launch_bounds__(NUM_THREADS, (N_MMA <= 64 ? 2 : 1))
ts_gemm_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_aux,
               const __grid_constant__ CUtensorMap tm_x, const Params p) {
  using C = Cfg<N_MMA>;
  constexpr int S = C::STAGES;
  constexpr int T = C::A_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * C::STAGE_BYTES);
  uint64_t* wfull = bars;             // [S]
  uint64_t* xfull = wfull + S;        // [S]
  uint64_t* sempty = xfull + S;       // [S]   4 dequant warps (the chunk's warpgroup) + MMA commit
  uint64_t* afull = sempty + S;       // [T]   4 dequant warps
  uint64_t* aempty = afull + T;       // [T]   MMA commit
  uint64_t* dfull = aempty + T;       // [2]
  uint64_t* dempty = dfull + 2;       // [2]   4 epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dempty + 2);
  uint32_t* flag_slot = tmem_slot + 1;

  // role (virtual) warp id; shifting by two warpgroups keeps the TMEM lane quarter (warp & 3) intact
  const int warp = ((threadIdx.x >> 5) + ((p.flags & 1) ? 8 : 0)) & 15, lane = threadIdx.x & 31;
  const bool nap = (p.flags & 2) != 0;
  const int G = gridDim.x, b = blockIdx.x;
  const long long t_entry = TL ? clock64() : 0;
  auto stamp = [&](int e) {
    if (TL && p.timeline) p.timeline[(size_t)b * 8 + e] = (unsigned long long)(clock64() - t_entry);
  };
  // fine-grained stamps of units 8..11 of CTA 0 (chain latencies), stored after the per-CTA table
  auto stamp2 = [&](int i, int e) {
    if (TL && p.timeline && b == 0 && i >= 8 && i < 12) p.timeline[148 * 8 + (i - 8) * 8 + e] = (unsigned long long)(clock64() - t_entry);
  };
  if (TL && p.timeline && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    p.timeline[(size_t)b * 8] = gt;
  }
  const long long U = (long long)p.n_tiles * p.m_blocks * p.KT;
  const int u0 = unit_begin(b, U, G), u1 = unit_begin(b + 1, U, G);
  const int nunits = u1 - u0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&wfull[i], 1);
      mbar_init(&xfull[i], 1);
      mbar_init(&sempty[i], 5);
    }
    for (int i = 0; i < T; ++i) {
      mbar_init(&afull[i], 4);
      mbar_init(&aempty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&dfull[i], 1);
      mbar_init(&dempty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == TMA_WARP && lane == 0) {
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_aux);
    tma_prefetch_desc(&tm_x);
  }
  if (warp == MMA_WARP) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  if (threadIdx.x == 0) stamp(1);

  // unit i of this CTA -> (tile, kc); tiles are (m_blk, n_tile) pairs, n_tile fastest
  auto tile_of = [&](int i) { return (u0 + i) / p.KT; };
  auto kc_of = [&](int i) { return (u0 + i) % p.KT; };

  if (warp < DEQ_WARPS) {
    // ------------------------------------------------------------ dequant warpgroups (0: even, 1: odd chunks)
    reg_inc<REGS_DEQ>();
    const int wg = warp >> 2, q4 = warp & 3;
    const int r = q4 * 32 + lane;  // weight row of the tile == TMEM lane
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    for (int i = wg; i < nunits; i += 2) {
      const int s = i % S, t = i % T;
      const uint32_t st = smem_u32(smem + (size_t)s * C::STAGE_BYTES);
      mbar_wait(&wfull[s], (i / S) & 1);
      if (i == 0 && warp == 0 && lane == 0) stamp(3);
      if ((warp & 3) == 0 && lane == 0) stamp2(i, 0);
      uint32_t out[64];
      Fmt::dequant(p, st, st + W_BYTES + C::X_BYTES, r, out);
      if ((warp & 3) == 0 && lane == 0) stamp2(i, 1);
      mbar_wait(&aempty[t], ((i / T) & 1) ^ 1);
      tc_fence_after();
      if ((warp & 3) == 0 && lane == 0) stamp2(i, 2);
      const uint32_t a_t = lane_taddr + C::A_COL0 + t * A_COLS;
      tmem_st_x32(a_t, out);
      tmem_st_x32(a_t + 32, out + 32);
      tc_wait_st();
      if ((warp & 3) == 0 && lane == 0) stamp2(i, 3);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&afull[t]);
        mbar_arrive(&sempty[s]);
      }
      if ((warp & 3) == 0 && lane == 0) stamp2(i, 4);
    }
  } else if (warp >= TMA_WARP) {
    reg_dec<REGS_OTHER>();
    if (warp == TMA_WARP) {
      if (lane == 0 && nunits > 0) {
        const uint64_t pol_w = policy_evict_first();
        const uint64_t pol_x = policy_evict_last();
        auto issue_w = [&](int i) {
          const int s = i % S, tile = tile_of(i);
          uint8_t* st = smem + (size_t)s * C::STAGE_BYTES;
          mbar_expect_tx(&wfull[s], Fmt::w_tx_bytes(p));
          Fmt::issue_w(&tm_w, &tm_aux, p, st, st + W_BYTES + C::X_BYTES, &wfull[s], tile % p.n_tiles, kc_of(i), pol_w);
        };
        auto issue_x = [&](int i) {
          const int s = i % S, tile = tile_of(i);
          uint8_t* st = smem + (size_t)s * C::STAGE_BYTES + W_BYTES;
          const int m0 = (tile / p.n_tiles) * N_MMA, k0 = kc_of(i) * KCHUNK;
          mbar_expect_tx(&xfull[s], C::X_BYTES);
          tma_load_2d(st, &tm_x, &xfull[s], k0, m0, pol_x);
          tma_load_2d(st + N_MMA * 128, &tm_x, &xfull[s], k0 + 64, m0, pol_x);
        };
        const int pre = nunits < S ? nunits : S;
        for (int i = 0; i < pre; ++i) issue_w(i);  // weights never depend on the previous kernel
        pdl_wait();
        stamp(2);
        for (int i = 0; i < pre; ++i) issue_x(i);
        for (int i = S; i < nunits; ++i) {
          while (!mbar_try_wait(&sempty[i % S], ((i / S) & 1) ^ 1)) { if (nap) __nanosleep(64); }
          issue_w(i);
          issue_x(i);
        }
      }
    } else if (warp == MMA_WARP) {
      constexpr uint32_t idesc = make_idesc(1 /*f32*/, 1 /*bf16*/, 1 /*bf16*/, ROWS, N_MMA);
      int seg = 0;  // accumulator segment counter (one per tile touched)
      for (int i = 0; i < nunits; ++i) {
        const int s = i % S, t = i % T;
        const bool first = (i == 0) || (kc_of(i) == 0);
        const bool last = (i == nunits - 1) || (kc_of(i) == p.KT - 1);
        const int buf = seg & 1;
        if (first) {
          mbar_wait(&dempty[buf], ((seg >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
        }
        while (!mbar_try_wait(&xfull[s], (i / S) & 1)) { if (nap) __nanosleep(32); }
        while (!mbar_try_wait(&afull[t], (i / T) & 1)) { if (nap) __nanosleep(32); }
        tc_fence_after();
        if (lane == 0) {
          if (i == 0) stamp(4);
          stamp2(i, 5);
          const uint32_t xb = smem_u32(smem + (size_t)s * C::STAGE_BYTES + W_BYTES);
          mma_chunk_ts_f16(tmem_base + (buf ? C::D_COL1 : C::D_COL0), tmem_base + C::A_COL0 + t * A_COLS,
                           umma_desc_k_sw128(xb), umma_desc_k_sw128(xb + N_MMA * 128), idesc, first ? 0u : 1u);
          tc_commit(&aempty[t]);
          tc_commit(&sempty[s]);
          if (last) tc_commit(&dfull[buf]);
          stamp2(i, 6);
          if (i == nunits - 1) stamp(5);
        }
        __syncwarp();
        if (last) ++seg;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (8..11)
    reg_dec<REGS_OTHER>();
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    pdl_wait();
    int seg = 0;
    int i = 0;
    while (i < nunits) {
      const int tile = tile_of(i);
      const int kc_first = kc_of(i);
      int cnt = p.KT - kc_first;
      if (cnt > nunits - i) cnt = nunits - i;
      const int buf = seg & 1;
      while (!mbar_try_wait(&dfull[buf], (seg >> 1) & 1)) __nanosleep(200);  // long wait: do not steal issue slots
      tc_fence_after();
      if (i + cnt >= nunits && (warp == EPI_WARP0 && lane == 0)) stamp(6);
      const int n_tile = tile % p.n_tiles, m_blk = tile / p.n_tiles;
      const int n = n_tile * ROWS + r, m0 = m_blk * N_MMA;
      const uint32_t d_t = lane_taddr + (buf ? C::D_COL1 : C::D_COL0);
      const float bias = (p.bias && n < p.N_out) ? __bfloat162float(p.bias[n]) : 0.f;
      const float osc = p.out_scale ? *p.out_scale : 1.f;
      if (cnt == p.KT) {
        // the whole K range of this tile was ours: straight to the output
#pragma unroll 1
        for (int j = 0; j < N_MMA; j += 8) {
          uint32_t rr[8];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7])
                       : "r"(d_t + j)
                       : "memory");
          tc_wait_ld();
          if (n < p.N_out) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int m = m0 + j + q;
              if (m < p.M) {
                float v = __uint_as_float(rr[q]);
                if (p.row_scale) v *= p.row_scale[m];
                p.y[(size_t)m * p.N_out + n] = __float2bfloat16_rn(v * osc + bias);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&dempty[buf]);
      } else {
        // tile shared with other CTAs: publish the partial, then see whether we complete the tile
        const int which = (u0 / p.KT == tile) ? 0 : 1;
        float* slot = p.ws_partial + ((size_t)b * 2 + which) * (N_MMA * ROWS);
#pragma unroll 1
        for (int j = 0; j < N_MMA; j += 8) {
          uint32_t rr[8];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7])
                       : "r"(d_t + j)
                       : "memory");
          tc_wait_ld();
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (m0 + j + q < p.M) __stcg(&slot[(j + q) * ROWS + r], __uint_as_float(rr[q]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&dempty[buf]);
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if ((warp == EPI_WARP0 && lane == 0)) {
          const unsigned prev = atomicAdd(&p.ws_sem[tile], (unsigned)cnt);
          *flag_slot = (prev + (unsigned)cnt == (unsigned)p.KT) ? 1u : 0u;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const bool finish = (*flag_slot != 0);
        asm volatile("bar.sync 1, 128;" ::: "memory");  // flag_slot is rewritten by the next segment
        if (finish) {
          __threadfence();
          const int b_first = cta_of_unit(tile * p.KT, U, G);
          const int b_last = cta_of_unit(tile * p.KT + p.KT - 1, U, G);
          const bool first_is_tail = unit_begin(b_first, U, G) < tile * p.KT;
          if ((warp == EPI_WARP0 && lane == 0)) p.ws_sem[tile] = 0;  // restore for the next launch
          if (n < p.N_out) {
            // fixed CTA order => bit-reproducible whoever finishes; 8 independent loads in flight per CTA slot
#pragma unroll 1
            for (int j0 = 0; j0 < N_MMA; j0 += 8) {
              if (m0 + j0 >= p.M) break;
              float v[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) v[q] = 0.f;
              for (int bb = b_first; bb <= b_last; ++bb) {
                // only the first contributor can have started in an earlier tile (then this is its tail slot)
                const int wh = (bb == b_first && first_is_tail) ? 1 : 0;
                const float* src = p.ws_partial + ((size_t)bb * 2 + wh) * (N_MMA * ROWS) + (size_t)j0 * ROWS + r;
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] += __ldcg(src + q * ROWS);
              }
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const int m = m0 + j0 + q;
                if (m < p.M) {
                  float t = v[q];
                  if (p.row_scale) t *= p.row_scale[m];
                  p.y[(size_t)m * p.N_out + n] = __float2bfloat16_rn(t * osc + bias);
                }
              }
            }
          }
        }
      }
      i += cnt;
      ++seg;
    }
    if ((warp == EPI_WARP0 && lane == 0)) stamp(7);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

}  // namespace tsg
}  // namespace ao
