// Persistent "dequantise-into-TMEM" decode GEMM for sub-byte weights with bf16 activations (sm_100a).
//
//   Y[M,N] = X[M,K] * W^[N,K]^T (+bias),   W^ produced in-kernel from 4-bit weights + group scales
//
// Used by int4_linear.cu (tinygemm tile-packed int4, W^ = bf16((q-8)s+z)) and
// nvfp4_weight_linear.cu (e2m1 * e4m3 block scale).  Structure:
//   * swap-AB: 128 weight rows = UMMA M, tokens = UMMA N (16..128), fp32 accumulator in TMEM
//   * persistent stream-K: the (n-tile, 128-k chunk) units of the whole GEMM are split EVENLY over the CTAs, so every
//     SM streams the same number of bytes whatever N/K are.  Grid (launchers): two CTAs per SM when a CTA would get
//     fewer than 16 chunks, else one (room for the next linear's CTA to become resident under this one, PDL)
//   * 16 warps in four warpgroups, 64 registers per thread, 256 TMEM columns, ~105 KB smem (two CTAs fit per SM):
//       WG0, WG1 (warps 0-7): dequant, alternating chunks: ld.shared (conflict-free through the TMA swizzle) ->
//                  unpack/scale in bf16x2, one 64-k half row (32 registers) at a time -> tcgen05.st of the bf16
//                  A operand into one of 3 TMEM A stages
//       WG2 (warps 8-11): epilogue: tcgen05.ld of a finished accumulator (double-buffered in TMEM: overlaps the
//                  next tile's MMAs), split-tile reduction through an fp32 workspace, bias, store
//       WG3: warp 12 weight TMA producer (never waits for the previous kernel), warp 14 activation TMA producer
//                  (after griddepcontrol.wait), warp 13 MMA issuer (warp 15: second issuer when Cfg::NI == 2)
//   * single-thread roles are WARP-UNIFORM loops with only the tcgen05 / TMA / mbarrier instruction under elect.sync
//     (warp index through __shfl_sync so the compiler knows it is uniform): with a loop under `lane == 0` ptxas wraps
//     every UTCHMMA / UTMALDG in an elect-broadcast loop and one thread issues an MMA only every ~52 cycles instead
//     of ~20, the tensor pipe's own floor for M128 N16 K16 (scripts/mma_microbench7.cu)
//   * the issuer does ONE wait and ONE commit per chunk: the activation slot of a chunk shares the index and the
//     barriers of the chunk's TMEM A stage (afull = 4 dequant-warp arrivals + the activation tile's TMA transaction
//     bytes; one tcgen05.commit on aempty frees both).  A-stage barriers are PAIRS per stage (see below)
//   * tiles split across CTAs are reduced deterministically: every CTA writes its partial, bumps the tile's unit
//     counter, and whoever completes the count sums the partials in CTA order (bit-reproducible run to run); the
//     CTA's LAST such reduction is on the kernel's critical path and is shared by all four warpgroups
//   * PDL: griddepcontrol.launch_dependents at start; only activations / outputs / workspace wait
//     (griddepcontrol.wait) for the previous kernel.
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
constexpr int DEQ_WARPS = 8, EPI_WARP0 = 8, TMA_WARP = 12, MMA_WARP = 13, XTMA_WARP = 14, MMA_WARP1 = 15;
constexpr int WSTAGE_BYTES = W_BYTES + AUX_BYTES;  // one weight stage: packed nibbles + scales
constexpr int NUM_THREADS = 16 * 32;

// DBUF = accumulator buffers per issuer (2: the epilogue of a tile overlaps the next tile's MMAs)
template <int N_MMA, int DBUF = 2>
struct Cfg {
  static constexpr int X_BYTES = 2 * N_MMA * 128;   // activation tile of one chunk (two 64-k swizzle atoms)
  // MMA issuer warps, one private accumulator set each.  One issuer reaches the tensor pipe's own limit (20 cycles
  // per M128 N16 K16 MMA) once its operands live in uniform registers; the two-issuer path (NI = 2, even / odd
  // chunks, accumulators added by the epilogue in a fixed order) is kept for experiments.
  static constexpr int NI = 1;
  static constexpr int TMEM_COLS = N_MMA <= 64 ? 256 : 512;
  static constexpr int D_COLS = NI * DBUF * N_MMA;
  static constexpr int A_COL0 = D_COLS <= 64 ? 64 : (D_COLS <= 128 ? 128 : 256);
  static constexpr int A_STAGES = (TMEM_COLS - A_COL0) / A_COLS;  // 3, 2 or 4; also the depth of the activation ring
  static constexpr int BUDGET = N_MMA <= 64 ? 104 * 1024 : 172 * 1024;
  // weight stages (8, 8, 6 or 4): even, so that a stage is always consumed by the same dequant warpgroup and
  // every waiter of a barrier observes each of its phases
  static constexpr int STAGES = ((BUDGET - A_STAGES * X_BYTES) / WSTAGE_BYTES) & ~1;
  static constexpr int X_OFF = STAGES * WSTAGE_BYTES;
  static constexpr int BAR_OFF = X_OFF + A_STAGES * X_BYTES;
  static constexpr size_t SMEM_BYTES = (size_t)BAR_OFF + 1024 + 1024;
  __host__ __device__ static constexpr int d_col(int issuer, int buf) { return (issuer * DBUF + buf) * N_MMA; }
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
  int flags;  // bring-up switches (AO_B200_TS_FLAGS, results are garbage): 1 = skip dequant arithmetic + TMEM stores, 2 = skip MMAs
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
//   static void dequant_half(p, w smem, aux smem, row r, half h, out[32])   (128 threads; out[c] = bf16x2 of
//                                                                            k = 64h + 2c, 64h + 2c + 1)
// TL = true compiles the per-CTA phase-timestamp instrumentation in (bring-up builds only).
template <class Fmt, int N_MMA, bool TL = false, int DBUF = 2>
__global__ void __launch_bounds__(NUM_THREADS, (N_MMA <= 64 ? 2 : 1))
ts_gemm_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_aux,
               const __grid_constant__ CUtensorMap tm_x, const Params p) {
  using C = Cfg<N_MMA, DBUF>;
  constexpr int NI = C::NI;
  constexpr int S = C::STAGES;
  constexpr int T = C::A_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
  uint64_t* wfull = bars;             // [S]   weight TMA transaction
  uint64_t* sempty = wfull + S;       // [S]   4 dequant warps (the chunk's warpgroup) have the weights in registers
  // A-stage barriers come in PAIRS per stage (use k = chunk / T of the stage goes to barrier k & 1, phase
  // k >> 1): with T odd consecutive uses of a stage belong to different warpgroups / issuers, and a parity
  // wait by a party that skips every other phase would alias; per pair every waiter sees every phase.
  uint64_t* afull = sempty + S;       // [T][2] 4 dequant warps (A stage stored) + activation TMA (arrive.expect_tx)
  uint64_t* aempty = afull + 2 * T;   // [T][2] MMA commit: A stage and activation slot both free
  uint64_t* dfull = aempty + 2 * T;   // [2]   one arrival per issuer and accumulator segment
  uint64_t* dempty = dfull + 2;       // [2]   4 epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dempty + 2);
  uint32_t* flag_slot = tmem_slot + 1;
  uint32_t* coop_slot = tmem_slot + 2;

  // The warp index goes through a shuffle so that the compiler knows it is warp-uniform: the single-thread
  // roles below run as warp-uniform loops with only the tcgen05 / TMA / mbarrier instruction itself under
  // elect.sync.  With the whole loop under `lane == 0` instead, ptxas treats every operand as divergent and
  // wraps each UTCHMMA / UTMALDG in an elect-broadcast "waterfall" loop (52 instead of 20 cycles per MMA,
  // scripts/mma_microbench7.cu).
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int G = gridDim.x, b = blockIdx.x;
  const long long t_entry = TL ? clock64() : 0;
  auto stamp = [&](int e) {
    if (TL && p.timeline && b < 100) {   // absolute globaltimer ns: comparable across back-to-back kernels
      unsigned long long gt;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
      p.timeline[(size_t)b * 8 + e] = gt;
    }
  };
  // fine-grained stamps of units 8..11 of CTA 0 (chain latencies), stored after the per-CTA table
  auto stamp2 = [&](int i, int e) {
    if (TL && p.timeline && b == 0 && i >= 8 && i < 12) p.timeline[100 * 8 + (i - 8) * 8 + e] = (unsigned long long)(clock64() - t_entry);
  };
  if (TL && p.timeline && threadIdx.x == 0 && b < 100) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    p.timeline[(size_t)b * 8] = gt;
  }
  const long long U = (long long)p.n_tiles * p.m_blocks * p.KT;
  const int u0 = unit_begin(b, U, G), u1 = unit_begin(b + 1, U, G);
  const int nunits = u1 - u0;

  if (threadIdx.x == 0) {
    *coop_slot = 0;
    for (int i = 0; i < S; ++i) {
      mbar_init(&wfull[i], 1);
      mbar_init(&sempty[i], 4);
    }
    for (int i = 0; i < 2 * T; ++i) {
      mbar_init(&afull[i], 5);
      mbar_init(&aempty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&dfull[i], NI);
      mbar_init(&dempty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == TMA_WARP && lane == 0) {
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_aux);
  }
  if (warp == XTMA_WARP && lane == 0) tma_prefetch_desc(&tm_x);
  if (warp == MMA_WARP) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_launch_dependents();
  if (threadIdx.x == 0) stamp(1);

  // unit i of this CTA -> (tile, kc); tiles are (m_blk, n_tile) pairs, n_tile fastest
  auto tile_of = [&](int i) { return (u0 + i) / p.KT; };
  auto kc_of = [&](int i) { return (u0 + i) % p.KT; };

  // Split-tile reduction of one tile by this CTA (it completed the tile's unit counter): sums the contributors'
  // partials in CTA order -- fixed order, so bit-reproducible whoever finishes -- and writes output row r of the
  // column groups g = helper, helper + nhelp, ... (8 token columns each).  The gather is a chain of L2 round trips, so
  // what counts is loads in flight: two contributors x 8 columns per step and, for the CTA's last segment, all four
  // warpgroups sharing the column groups (dequant warps, producers and issuer are idle by then).
  auto finish_tile = [&](int tile, int r, int helper, int nhelp) {
    const int b_first = cta_of_unit(tile * p.KT, U, G);
    const int b_last = cta_of_unit(tile * p.KT + p.KT - 1, U, G);
    const bool first_is_tail = unit_begin(b_first, U, G) < tile * p.KT;
    const int n_tile = tile % p.n_tiles, m_blk = tile / p.n_tiles;
    const int n = n_tile * ROWS + r, m0 = m_blk * N_MMA;
    if (n >= p.N_out) return;
    const float bias = p.bias ? __bfloat162float(p.bias[n]) : 0.f;
    const float osc = p.out_scale ? *p.out_scale : 1.f;
    auto slot_of = [&](int bb) {
      // only the first contributor can have started in an earlier tile (then this is its tail slot)
      const int wh = (bb == b_first && first_is_tail) ? 1 : 0;
      return p.ws_partial + ((size_t)bb * 2 + wh) * (N_MMA * ROWS) + r;
    };
    if (p.M - m0 == 1) {
      // decode, one token column: every contributor's value in flight at once (one L2 round trip per 8)
      if (helper != 0) return;
      float acc = 0.f;
#pragma unroll 1
      for (int bb = b_first; bb <= b_last; bb += 8) {
        float t[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) t[c] = (bb + c <= b_last) ? __ldcg(slot_of(bb + c)) : 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) acc += t[c];
      }
      if (p.row_scale) acc *= p.row_scale[m0];
      p.y[(size_t)m0 * p.N_out + n] = __float2bfloat16_rn(acc * osc + bias);
      return;
    }
#pragma unroll 1
    for (int j0 = helper * 8; j0 < N_MMA; j0 += nhelp * 8) {
      if (m0 + j0 >= p.M) break;
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = 0.f;
      int bb = b_first;
#pragma unroll 1
      for (; bb + 1 <= b_last; bb += 2) {
        const float* s0 = slot_of(bb) + (size_t)j0 * ROWS;
        const float* s1 = slot_of(bb + 1) + (size_t)j0 * ROWS;
        float t0[8], t1[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) t0[q] = (m0 + j0 + q < p.M) ? __ldcg(s0 + q * ROWS) : 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) t1[q] = (m0 + j0 + q < p.M) ? __ldcg(s1 + q * ROWS) : 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = (v[q] + t0[q]) + t1[q];
      }
      if (bb <= b_last) {
        const float* s0 = slot_of(bb) + (size_t)j0 * ROWS;
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] += (m0 + j0 + q < p.M) ? __ldcg(s0 + q * ROWS) : 0.f;
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
  };
  // hand-over of the CTA's last split-tile reduction to all four warpgroups: coop_slot = tile + 1, or 0
  // (only when several column groups exist to share: with <= 8 token columns the hand-over costs more than it saves)
  const bool use_coop = p.M > 8;
  auto coop_barrier = [&]() { asm volatile("bar.sync 2, 512;" ::: "memory"); };

  if (warp < DEQ_WARPS) {
    // ------------------------------------------------------------ dequant warpgroups (0: even, 1: odd chunks)
    const int wg = warp >> 2, q4 = warp & 3;
    const int r = q4 * 32 + lane;  // weight row of the tile == TMEM lane
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    // ring positions of chunk i = wg, wg + 2, ... kept incrementally (the integer pipe is the bottleneck here)
    int s = wg, sph = 0, t = wg % T, k = wg / T;
    for (int i = wg; i < nunits; i += 2) {
      const uint32_t st = smem_u32(smem + (size_t)s * WSTAGE_BYTES);
      const uint32_t a_t = lane_taddr + C::A_COL0 + t * A_COLS;
      mbar_wait(&wfull[s], sph);
      if (i == 0 && warp == 0 && lane == 0) stamp(3);
      if ((warp & 3) == 0 && lane == 0) stamp2(i, 0);
      // the row in two 64-k halves (32 registers of output each): half 0 is computed before the A stage is
      // known to be free, so the wait overlaps its arithmetic
      uint32_t out[32];
      if (p.flags & 1) {   // bring-up: skip the dequant arithmetic and the TMEM stores (timing experiments only)
        if (k >= 1) mbar_wait(&aempty[t * 2 + ((k - 1) & 1)], ((k - 1) >> 1) & 1);
        __syncwarp();
        if (elect_one()) { mbar_arrive(&sempty[s]); mbar_arrive(&afull[t * 2 + (k & 1)]); }
        s += 2;
        if (s >= S) { s -= S; sph ^= 1; }
        t += 2;
        if (t >= T) { t -= T; ++k; }
        continue;
      }
      Fmt::dequant_half(p, st, st + W_BYTES, r, 0, out);
      if ((warp & 3) == 0 && lane == 0) stamp2(i, 1);
      if (k >= 1) mbar_wait(&aempty[t * 2 + ((k - 1) & 1)], ((k - 1) >> 1) & 1);  // MMAs of chunk i - T are done
      tc_fence_after();
      if ((warp & 3) == 0 && lane == 0) stamp2(i, 2);
      tmem_st_x32(a_t, out);   // source registers are consumed at issue: no tcgen05.wait::st before reusing them
      Fmt::dequant_half(p, st, st + W_BYTES, r, 1, out);
      __syncwarp();
      if (elect_one()) mbar_arrive(&sempty[s]);  // weights are in registers: the stage can be refilled
      tmem_st_x32(a_t + 32, out);
      tc_wait_st();
      if ((warp & 3) == 0 && lane == 0) stamp2(i, 3);
      tc_fence_before();
      __syncwarp();
      if (elect_one()) mbar_arrive(&afull[t * 2 + (k & 1)]);
      if ((warp & 3) == 0 && lane == 0) stamp2(i, 4);
      s += 2;
      if (s >= S) { s -= S; sph ^= 1; }
      t += 2;
      if (t >= T) { t -= T; ++k; }
    }
    // help with the CTA's last split-tile reduction, if this CTA turns out to be the one completing that tile
    if (use_coop) coop_barrier();
    const uint32_t ct = use_coop ? *coop_slot : 0u;
    if (ct) {
      __threadfence();
      finish_tile((int)ct - 1, r, 1 + wg, 4);
    }
  } else if (warp >= TMA_WARP) {
    if (warp == TMA_WARP) {
      // ---------------------------------------------------------- weight producer.  Weights never depend on the
      // previous kernel: no griddepcontrol.wait on this path.
      const uint64_t pol_w = policy_evict_first();
      int s = 0, sph = 0, kc = kc_of(0), n_tile = tile_of(0) % p.n_tiles;
      for (int i = 0; i < nunits; ++i) {
        if (i >= S) mbar_wait(&sempty[s], sph ^ 1);
        if (elect_one()) {
          uint8_t* st = smem + (size_t)s * WSTAGE_BYTES;
          mbar_expect_tx(&wfull[s], Fmt::w_tx_bytes(p));
          Fmt::issue_w(&tm_w, &tm_aux, p, st, st + W_BYTES, &wfull[s], n_tile, kc, pol_w);
        }
        __syncwarp();
        if (++s == S) { s = 0; sph ^= 1; }
        if (++kc == p.KT) { kc = 0; if (++n_tile == p.n_tiles) n_tile = 0; }
      }
    } else if (warp == XTMA_WARP) {
      // ---------------------------------------------------------- activation producer
      if (nunits > 0) {
        const uint64_t pol_x = policy_evict_last();
        pdl_wait();   // activations are the previous kernel's output
        if (lane == 0) stamp(2);
        int t = 0, k = 0, kc = kc_of(0), tile = tile_of(0);
        for (int i = 0; i < nunits; ++i) {
          uint64_t* full = &afull[t * 2 + (k & 1)];
          if (k >= 1) mbar_wait(&aempty[t * 2 + ((k - 1) & 1)], ((k - 1) >> 1) & 1);
          if (elect_one()) {
            uint8_t* xs = smem + C::X_OFF + (size_t)t * C::X_BYTES;
            const int m0 = (tile / p.n_tiles) * N_MMA, k0 = kc * KCHUNK;
            mbar_expect_tx(full, C::X_BYTES);
            tma_load_2d(xs, &tm_x, full, k0, m0, pol_x);
            tma_load_2d(xs + N_MMA * 128, &tm_x, full, k0 + 64, m0, pol_x);
          }
          __syncwarp();
          if (++t == T) { t = 0; ++k; }
          if (++kc == p.KT) { kc = 0; ++tile; }
        }
      }
    } else if (warp == MMA_WARP || (NI == 2 && warp == MMA_WARP1)) {
      // ---------------------------------------------------------- MMA issuers: chunk c belongs to issuer c % NI
      constexpr uint32_t idesc = make_idesc(1 /*f32*/, 1 /*bf16*/, 1 /*bf16*/, ROWS, N_MMA);
      const int j = (warp == MMA_WARP) ? 0 : 1;
      const uint32_t x0 = smem_u32(smem + C::X_OFF);
      int seg = 0, i = 0, kc0 = kc_of(0);
      while (i < nunits) {
        int cnt = p.KT - kc0;   // units of this accumulator segment
        if (cnt > nunits - i) cnt = nunits - i;
        const int buf = seg % DBUF, ph = (seg / DBUF) & 1;
        // every issuer passes through every segment's dempty/dfull phase, chunks or not (keeps the phases in step)
        mbar_wait(&dempty[buf], ph ^ 1);
        const uint32_t d_t = tmem_base + C::d_col(j, buf);
        int c = i + ((j - i) & (NI - 1));
        int t = c % T, k = c / T;
        uint32_t acc = 0;   // first MMA of the segment overwrites the accumulator
        for (; c < i + cnt; c += NI) {
          mbar_wait(&afull[t * 2 + (k & 1)], (k >> 1) & 1);
          tc_fence_after();
          if (elect_one()) {
            if (c == 0) stamp(4);
            stamp2(c, 5);
            const uint32_t xb = x0 + t * C::X_BYTES;
            const uint32_t a_t = tmem_base + C::A_COL0 + t * A_COLS;
            const uint64_t b_lo = umma_desc_k_sw128(xb), b_hi = umma_desc_k_sw128(xb + N_MMA * 128);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              if (!(p.flags & 2))   // bring-up: flag 2 skips the MMAs
                mma_ts_f16(d_t, a_t + kk * 8, (kk < 4 ? b_lo : b_hi) + (uint64_t)((kk & 3) * 2), idesc, (kk == 0) ? acc : 1u);
            tc_commit(&aempty[t * 2 + (k & 1)]);
            stamp2(c, 6);
            if (c == nunits - 1) stamp(5);
          }
          __syncwarp();
          acc = 1u;
          t += NI;
          if (t >= T) { t -= T; ++k; }
        }
        if (elect_one()) {
          if (acc) tc_commit(&dfull[buf]);   // all of this issuer's MMAs of the segment have completed
          else mbar_arrive(&dfull[buf]);     // no chunk of this segment was ours
        }
        __syncwarp();
        i += cnt;
        ++seg;
        kc0 = 0;
      }
    }
    // fourth helper warpgroup of the CTA's last split-tile reduction (producers and issuer are done by then)
    if (use_coop) {
      coop_barrier();
      const uint32_t ct = *coop_slot;
      if (ct) {
        __threadfence();
        finish_tile((int)ct - 1, (warp & 3) * 32 + lane, 3, 4);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (8..11)
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    pdl_wait();
    int seg = 0;
    int i = 0;
    bool coop = false;
    while (i < nunits) {
      const int tile = tile_of(i);
      const int kc_first = kc_of(i);
      int cnt = p.KT - kc_first;
      if (cnt > nunits - i) cnt = nunits - i;
      const int buf = seg % DBUF;
      while (!mbar_try_wait(&dfull[buf], (seg / DBUF) & 1)) __nanosleep(64);  // long wait: do not steal issue slots
      tc_fence_after();
      if (i + cnt >= nunits && (warp == EPI_WARP0 && lane == 0)) stamp(6);
      const int n_tile = tile % p.n_tiles, m_blk = tile / p.n_tiles;
      const int n = n_tile * ROWS + r, m0 = m_blk * N_MMA;
      // which issuers contributed to this segment (chunk c -> issuer c % NI); their accumulators add in fixed order
      const bool has0 = (NI == 1) || cnt >= 2 || (i & 1) == 0;
      const bool has1 = (NI == 2) && (cnt >= 2 || (i & 1) == 1);
      const uint32_t d_t0 = lane_taddr + C::d_col(0, buf);
      const uint32_t d_t1 = lane_taddr + C::d_col(NI - 1, buf);
      auto load8 = [&](int j8, float* v) {
        uint32_t ra[8], rb[8];
        if (has0)
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(ra[0]), "=r"(ra[1]), "=r"(ra[2]), "=r"(ra[3]), "=r"(ra[4]), "=r"(ra[5]), "=r"(ra[6]), "=r"(ra[7])
                       : "r"(d_t0 + j8)
                       : "memory");
        if (has1)
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(rb[0]), "=r"(rb[1]), "=r"(rb[2]), "=r"(rb[3]), "=r"(rb[4]), "=r"(rb[5]), "=r"(rb[6]), "=r"(rb[7])
                       : "r"(d_t1 + j8)
                       : "memory");
        tc_wait_ld();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float a = has0 ? __uint_as_float(ra[q]) : 0.f;
          v[q] = has1 ? (has0 ? a + __uint_as_float(rb[q]) : __uint_as_float(rb[q])) : a;
        }
      };
      const float bias = (p.bias && n < p.N_out) ? __bfloat162float(p.bias[n]) : 0.f;
      const float osc = p.out_scale ? *p.out_scale : 1.f;
      if (cnt == p.KT) {
        // the whole K range of this tile was ours: straight to the output
#pragma unroll 1
        for (int j = 0; j < N_MMA; j += 8) {
          float rr[8];
          load8(j, rr);
          if (n < p.N_out) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int m = m0 + j + q;
              if (m < p.M) {
                float v = rr[q];
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
          float rr[8];
          load8(j, rr);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (m0 + j + q < p.M) __stcg(&slot[(j + q) * ROWS + r], rr[q]);
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
          if ((warp == EPI_WARP0 && lane == 0)) p.ws_sem[tile] = 0;  // restore for the next launch
          if (use_coop && i + cnt >= nunits) {
            // the CTA's last segment: the reduction is on the kernel's critical path, share it (see finish_tile)
            if (warp == EPI_WARP0 && lane == 0) *coop_slot = (uint32_t)tile + 1u;
            coop = true;
          } else {
            finish_tile(tile, r, 0, 1);
          }
        }
      }
      i += cnt;
      ++seg;
    }
    if (use_coop) coop_barrier();
    if (coop) finish_tile((int)*coop_slot - 1, r, 0, 4);
    if ((warp == EPI_WARP0 && lane == 0)) stamp(7);
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

}  // namespace tsg
}  // namespace ao
