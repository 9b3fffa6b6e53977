// Prefill-shaped (M > 128 tokens) "dequantise-into-TMEM" GEMM for sub-byte weights with bf16 activations (sm_100a).
//
//   Y[M,N] = X[M,K] * W^[N,K]^T (+bias),   W^ produced in-kernel from 4-bit weights + group scales
//
// Same TS-mode pipeline and format policies as ts_gemm.cuh (the decode kernel), re-proportioned for the
// tensor-core-bound regime:
//   * tile = 128 weight rows (UMMA M) x 256 TOKENS (UMMA N): a weight chunk is fetched and dequantised ONCE per
//     256 tokens (the decode kernel re-streams and re-dequantises the weights for every 128-token block), and one
//     tcgen05.mma covers 128 x 256 x 16: 8 MMAs = 1024 tensor-pipe cycles per chunk against ~600 cycles of dequant
//     work per chunk on the three warpgroups, so the tensor pipe, not the integer pipe, is the bound
//   * TMEM: accumulator 128 lanes x 256 fp32 columns + 4 bf16 A stages of 64 columns = all 512 columns.  The
//     accumulator is single-buffered; a segment's epilogue is shared by the three dequant warpgroups (every third
//     group of 8 token columns each) and DEFERRED by one chunk per warpgroup, so three chunks of the next segment are
//     already dequantised when the accumulator is handed back and the MMAs restart at once
//   * shared memory (1 CTA per SM): 6 weight stages (10 KB) + 5 activation slots of HALF a chunk (64 k x 256
//     tokens = 32 KB: what one TMA box / one 128-byte swizzle atom holds).  Activation slots have their own barrier
//     pair (the decode kernel ties a chunk's activation tile to its A stage)
//   * persistent stream-K with the owner-gather fix-up of streamk.cuh (tiles x K chunks split evenly over the
//     SMs: no wave quantisation at M = 512, where a Llama projection has fewer 128 x 256 tiles than the GPU has SMs)
//   * warps 0-11 dequant + epilogue, 12 weight TMA, 13 MMA issuer, 14 activation TMA, 15 idle.
// Numerics are those of the decode kernel: bf16 W^ bit-identical to the reference dequant, fp32 accumulation over K in
// TMEM, split tiles summed in CTA (= k) order: deterministic.
#pragma once
#include <cuda_bf16.h>

#include "common.h"
#include "ptx.cuh"
#include "streamk.cuh"
#include "ts_gemm.cuh"

namespace ao {
namespace tsp {

using streamk::ROWS;
using tsg::A_COLS;
using tsg::KCHUNK;
using tsg::W_BYTES;
using tsg::WSTAGE_BYTES;

constexpr int N_TOK = 256;                 // tokens per tile = UMMA N
constexpr int S = 6;                       // weight stages: a multiple of the 3 dequant warpgroups, so a stage always
                                           // belongs to the same warpgroup and its parity waits cannot alias (ts_gemm.cuh)
constexpr int T = 4;                       // TMEM A stages
constexpr int XS = 5;                      // activation half-chunk slots
constexpr int XH_BYTES = N_TOK * 128;      // 64 k x 256 tokens of bf16 = 32 KiB
constexpr int X_OFF = S * WSTAGE_BYTES;
constexpr int BAR_OFF = X_OFF + XS * XH_BYTES;
constexpr size_t SMEM_BYTES = (size_t)BAR_OFF + 1024 + 1024;
constexpr int TMEM_COLS = 512, A_COL0 = 256;
constexpr int DEQ_WGS = 3, DEQ_WARPS = 12, TMA_WARP = 12, MMA_WARP = 13, XTMA_WARP = 14;
constexpr int NUM_THREADS = 16 * 32;
static_assert(SMEM_BYTES <= 227 * 1024, "prefill kernel: shared memory budget");
static_assert(S % DEQ_WGS == 0, "a weight stage must always belong to the same dequant warpgroup");

__device__ __forceinline__ uint64_t policy_evict_normal() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}

template <class Fmt>
__global__ void __launch_bounds__(NUM_THREADS, 1)
ts_prefill_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_aux,
                  const __grid_constant__ CUtensorMap tm_x, const tsg::Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* wfull = bars;            // [S]  weight TMA transaction
  uint64_t* wempty = wfull + S;      // [S]  4 dequant warps have the stage's rows in registers
  uint64_t* afull = wempty + S;      // [T]  4 dequant warps stored the bf16 A stage
  uint64_t* aempty = afull + T;      // [T]  MMA commit: the chunk's 8 MMAs are done
  uint64_t* xfull = aempty + T;      // [XS] activation TMA transaction
  uint64_t* xempty = xfull + XS;     // [XS] MMA commit: the half chunk's 4 MMAs are done
  uint64_t* dfull = xempty + XS;     // [1]  accumulator of a segment complete (phase = segment parity)
  uint64_t* dempty = dfull + 1;      // [1]  12 dequant warps have read their share of it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dempty + 1);

  // warp index through a shuffle: known warp-uniform, so the single-thread roles keep their operands in uniform
  // registers (ts_gemm.cuh)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int G = gridDim.x, b = blockIdx.x;
  const long long U = (long long)p.n_tiles * p.m_blocks * p.KT;
  const int u0 = streamk::unit_begin(b, U, G), u1 = streamk::unit_begin(b + 1, U, G);
  const int nunits = u1 - u0;
  const streamk::Walk walk(u0, nunits, p.KT);

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&wfull[i], 1);
      mbar_init(&wempty[i], 4);
    }
    for (int i = 0; i < T; ++i) {
      mbar_init(&afull[i], 4);
      mbar_init(&aempty[i], 1);
    }
    for (int i = 0; i < XS; ++i) {
      mbar_init(&xfull[i], 1);
      mbar_init(&xempty[i], 1);
    }
    mbar_init(dfull, 1);
    mbar_init(dempty, DEQ_WARPS);
    fence_barrier_init();
  }
  if (warp == TMA_WARP && lane == 0) {
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_aux);
  }
  if (warp == XTMA_WARP && lane == 0) tma_prefetch_desc(&tm_x);
  if (warp == MMA_WARP) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_launch_dependents();

  auto tile_of = [&](int i) { return (u0 + i) / p.KT; };
  auto kc_of = [&](int i) { return (u0 + i) % p.KT; };

  if (warp < DEQ_WARPS) {
    // ------------------------------------------------------------ dequant warpgroups (chunk i -> WG i % 3) + epilogues
    const int wg = warp >> 2, q4 = warp & 3;
    const int r = q4 * 32 + lane;  // weight row of the tile == TMEM lane
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    bool waited_prev = false;

    auto emit8 = [&](const float (&v)[8], int n, int m_first, float bias, float osc) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int m = m_first + q;
        if (m < p.M) {
          float t = v[q];
          if (p.row_scale) t *= p.row_scale[m];
          p.y[(size_t)m * p.N_out + n] = __float2bfloat16_rn(t * osc + bias);
        }
      }
    };

    // this warpgroup's share (token column groups wg, wg + 3, ... of 8) of segment `seg`
    auto epilogue = [&](int seg) {
      if (!waited_prev) { pdl_wait(); waited_prev = true; }   // outputs / workspace belong to the previous kernel
      const int tile = walk.seg_tile(seg);
      const int kind = walk.seg_kind(seg);
      mbar_wait(dfull, seg & 1);
      tc_fence_after();
      const int n_tile = tile % p.n_tiles, m_blk = tile / p.n_tiles;
      const int n = n_tile * ROWS + r, m0 = m_blk * N_TOK;
      const bool row_ok = n < p.N_out;
      if (kind == streamk::SEG_CONTRIB) {
        // publish the partial (column-major slot: coalesced across the 128 rows); flag after all three warpgroups
        float* slot = p.ws_partial + (size_t)b * (N_TOK * ROWS) + r;
#pragma unroll 1
        for (int j = wg * 8; j < N_TOK; j += DEQ_WGS * 8) {
          if (m0 + j >= p.M) break;
          float v[8];
          tsg::tmem_ld_x8(lane_taddr + j, v);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (m0 + j + q < p.M) __stcg(&slot[(j + q) * ROWS], v[q]);
        }
        asm volatile("bar.sync 1, 384;" ::: "memory");                       // every row / column stored (cta scope) ...
        if (warp == 0 && lane == 0) streamk::st_release_u32(p.ws_flag + b, 1u);   // ... then one gpu-scope release
      } else {
        const float bias = (p.bias && row_ok) ? __bfloat162float(p.bias[n]) : 0.f;
        const float osc = (p.out_scale ? (p.out_scale_per_row ? (row_ok ? p.out_scale[n] : 1.f) : *p.out_scale) : 1.f) *
                          __int_as_float((127 + p.acc_exp2) << 23);
        int n_oth = 0;
        const float* slot0 = nullptr;
        if (kind == streamk::SEG_OWNER) {
          // own partial (TMEM) + the partials of CTAs b+1 .. b_last in that order (= k order: deterministic)
          const int b_last = streamk::cta_of_unit((long long)tile * p.KT + p.KT - 1, U, G);
          n_oth = b_last - b;
          slot0 = p.ws_partial + (size_t)(b + 1) * (N_TOK * ROWS) + r;
          streamk::wait_flags(p.ws_flag + b + 1, n_oth, lane);
        }
#pragma unroll 1
        for (int j = wg * 8; j < N_TOK; j += DEQ_WGS * 8) {
          if (m0 + j >= p.M) break;
          float v[8];
          tsg::tmem_ld_x8(lane_taddr + j, v);
#pragma unroll 1
          for (int c0 = 0; c0 < n_oth; c0 += 2) {
            float t[2][8];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const float* sc = slot0 + (size_t)(c0 + c) * (N_TOK * ROWS) + (size_t)j * ROWS;
#pragma unroll
              for (int q = 0; q < 8; ++q)
                t[c][q] = (c0 + c < n_oth && row_ok && m0 + j + q < p.M) ? __ldcg(sc + q * ROWS) : 0.f;
            }
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
              for (int q = 0; q < 8; ++q) v[q] += t[c][q];
          }
          if (row_ok) emit8(v, n, m0 + j, bias, osc);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dempty);
    };
    int ep_seg = 0;
    auto seg_end = [&](int sg) { return walk.seg_begin(sg) + walk.seg_count(sg) - 1; };

    for (int i = wg; i < nunits; i += DEQ_WGS) {
      const int s = i % S, t = i & (T - 1);
      const uint32_t st = smem_u32(smem + (size_t)s * WSTAGE_BYTES);
      const uint32_t a_t = lane_taddr + A_COL0 + t * A_COLS;
      mbar_wait(&wfull[s], (i / S) & 1);
      typename Fmt::Raw raw;
      Fmt::load_row(p, st, st + W_BYTES, r, raw);
      // the row is in registers (enforced through the arrive's address, see ts_gemm.cuh): the stage can be refilled
      const uint32_t never = (Fmt::touch(raw) == 0x9E3779B9u) & (p.flags == 0x7fffffff);
      __syncwarp();
      if (elect_one()) mbar_arrive(&wempty[s] + never);
      uint32_t out[16];
      Fmt::dequant_quarter(p, raw, 0, out);
      if (i >= T) mbar_wait(&aempty[t], ((i / T) & 1) ^ 1);   // MMAs of chunk i - T are done: A stage t is free
      tc_fence_after();
      tmem_st_x16(a_t, out);
#pragma unroll
      for (int q = 1; q < 4; ++q) {
        Fmt::dequant_quarter(p, raw, q, out);
        tmem_st_x16(a_t + 16 * q, out);
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (elect_one()) mbar_arrive(&afull[t]);
      // segments that ended before this chunk: their accumulator is complete or about to be; this warpgroup's first
      // chunk after a segment's end is at most 3 past it, so the A stage it just filled never needed the accumulator
      while (ep_seg < walk.nseg && seg_end(ep_seg) < i) epilogue(ep_seg++);
    }
    while (ep_seg < walk.nseg) epilogue(ep_seg++);
  } else if (warp == TMA_WARP) {
    // ---------------------------------------------------------- weight producer (never waits for the previous kernel)
    const uint64_t pol_w = policy_evict_normal();   // the same weight tile is read again for the next 256 tokens
    int kc = kc_of(0), n_tile = tile_of(0) % p.n_tiles;
    for (int i = 0; i < nunits; ++i) {
      const int s = i % S;
      if (i >= S) mbar_wait(&wempty[s], ((i / S) & 1) ^ 1);
      if (elect_one()) {
        uint8_t* st = smem + (size_t)s * WSTAGE_BYTES;
        mbar_expect_tx(&wfull[s], Fmt::w_tx_bytes(p));
        Fmt::issue_w(&tm_w, &tm_aux, p, st, st + W_BYTES, &wfull[s], n_tile, kc, pol_w);
      }
      __syncwarp();
      if (++kc == p.KT) { kc = 0; if (++n_tile == p.n_tiles) n_tile = 0; }
    }
  } else if (warp == XTMA_WARP) {
    // ---------------------------------------------------------- activation producer: two half chunks per unit
    const uint64_t pol_x = policy_evict_last();
    pdl_wait();
    int c = 0, cph = 1, kc = kc_of(0), tile = tile_of(0);   // cph: parity of the slot's PREVIOUS use
    for (int h = 0; h < 2 * nunits; ++h) {
      if (h >= XS) mbar_wait(&xempty[c], cph);
      if (elect_one()) {
        uint8_t* xs = smem + X_OFF + (size_t)c * XH_BYTES;
        const int m0 = (tile / p.n_tiles) * N_TOK, k0 = kc * KCHUNK + (h & 1) * 64;
        mbar_expect_tx(&xfull[c], XH_BYTES);
        tma_load_2d(xs, &tm_x, &xfull[c], k0, m0, pol_x);
      }
      __syncwarp();
      if (++c == XS) { c = 0; cph ^= 1; }
      if (h & 1) { if (++kc == p.KT) { kc = 0; ++tile; } }
    }
  } else if (warp == MMA_WARP) {
    // ---------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = make_idesc(1 /*f32*/, 1 /*bf16*/, 1 /*bf16*/, ROWS, N_TOK);
    const uint32_t x0 = smem_u32(smem + X_OFF);
    pdl_wait();
    int seg = 0, seg_last_unit = walk.seg_count(0) - 1;
    int c = 0;
    uint32_t cph = 0, acc = 0;
    for (int i = 0; i < nunits; ++i) {
      const int t = i & (T - 1);
      mbar_wait(&afull[t], (i / T) & 1);
      tc_fence_after();
      const uint32_t a_t = tmem_base + A_COL0 + t * A_COLS;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        mbar_wait(&xfull[c], cph);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t bdesc = umma_desc_k_sw128(x0 + c * XH_BYTES);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            mma_ts_f16(tmem_base, a_t + (half * 4 + kk) * 8, bdesc + (uint64_t)(kk * 2), idesc, (half == 0 && kk == 0) ? acc : 1u);
          tc_commit(&xempty[c]);
          if (half == 1) tc_commit(&aempty[t]);
        }
        __syncwarp();
        if (++c == XS) { c = 0; cph ^= 1; }
      }
      acc = 1u;
      if (i == seg_last_unit) {
        if (elect_one()) tc_commit(dfull);
        __syncwarp();
        if (++seg < walk.nseg) {
          seg_last_unit += walk.seg_count(seg);
          mbar_wait(dempty, (seg - 1) & 1);   // the single accumulator has been read out
          tc_fence_after();
          acc = 0;
        }
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 0 && walk.seg_kind(walk.nseg - 1) == streamk::SEG_OWNER) {
    // every warpgroup has read the contributors' partials: re-arm their flags for the next launch
    const int b_last = streamk::cta_of_unit((long long)walk.seg_tile(walk.nseg - 1) * p.KT + p.KT - 1, U, G);
    for (int c = b + 1 + lane; c <= b_last; c += 32) p.ws_flag[c] = 0u;
  }
  if (warp == MMA_WARP) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// When to take this kernel instead of the decode kernel's 128-token blocks (M > 128): measured crossover on the
// Llama-3-8B projections (profiles/r02_call_q.log).  With fewer than ~50 chunks of 128 x 256 per SM the 128-token-block
// path is ahead (smaller tiles split more evenly, 16 KB instead of 128 KB of partials per split tile): e.g. 512 tokens,
// o-proj 23.2 vs 30.1 us, down-proj 49.6 vs 55.4 us; above it this kernel wins by up to 16 % (4096 tokens, gate|up 715 vs
// 832 us).
inline bool worth_it(int M, int N_out, int K) {
  if (M <= 128 || prefill_disabled()) return false;
  const long long units = (long long)ceil_div(N_out, ROWS) * ceil_div(M, N_TOK) * (K / KCHUNK);
  return units >= 50LL * sm_count();
}

// Grid + workspace carve-up: one CTA per SM, never fewer than 8 chunks per CTA.
inline int plan(tsg::Params& p, void* ws, size_t ws_bytes, const char* what, int* grid_out) {
  const long long units = (long long)p.n_tiles * p.m_blocks * p.KT;
  int grid = sm_count();
  const int min_units = 8;
  if (units / min_units < grid) grid = units / min_units > 0 ? (int)(units / min_units) : 1;
  const size_t need = streamk::WS_PARTIAL_OFF + (size_t)grid * N_TOK * ROWS * 4;
  if (!ws || ws_bytes < need || (size_t)grid * 4 > streamk::WS_FLAGS_BYTES)
    return fail(AO_ERR_WORKSPACE, "%s: workspace too small (%zu < %zu)", what, ws_bytes, need);
  p.ws_flag = reinterpret_cast<unsigned int*>(ws);
  p.ws_partial = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + streamk::WS_PARTIAL_OFF);
  p.flags = 0;
  p.producers = 1;
  p.prefetch = 0;
  p.timeline = nullptr;
  *grid_out = grid;
  return AO_OK;
}

}  // namespace tsp
}  // namespace ao
