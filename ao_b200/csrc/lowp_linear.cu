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
// Split-K across CTAs with a deterministic last-CTA reduction (int32 for int8: stays exact).
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace ao {
namespace lowp {

enum Kind { KIND_I8 = 0, KIND_F8 = 1, KIND_MXF8 = 2, KIND_NVF4 = 3 };

constexpr int ROWS = 128;
constexpr int KB = 128;  // bytes of K per stage per row (128 elems for 8-bit, 256 for fp4)
constexpr int A_BYTES = ROWS * KB;  // 16 KiB
constexpr int NUM_THREADS = 192;    // warp0 TMA, warp1 MMA, warps2-5 epilogue
constexpr int MAX_SPLITS = 32;

template <int KIND, int N_MMA>
struct Cfg {
  static constexpr bool BLOCK_SCALED = (KIND == KIND_MXF8 || KIND == KIND_NVF4);
  static constexpr int B_BYTES = N_MMA * KB;
  // scale tiles per stage: mxfp8 1 blocked tile (128 rows x 4 sf), nvfp4 4 tiles (16 sf / row)
  static constexpr int SF_TILES = KIND == KIND_MXF8 ? 1 : (KIND == KIND_NVF4 ? 4 : 0);
  static constexpr int SF_BYTES = SF_TILES * 512;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES + 2 * SF_BYTES + (BLOCK_SCALED ? (1024 - (2 * SF_BYTES) % 1024) % 1024 : 0);
  static constexpr int STAGES = (96 * 1024) / STAGE_BYTES < 3 ? 3 : ((96 * 1024) / STAGE_BYTES > 8 ? 8 : (96 * 1024) / STAGE_BYTES);
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 + 512;
  // TMEM: D at [0, N_MMA); scale factors after it (double-buffered per stage parity)
  static constexpr int SF_COLS = SF_TILES * 4;              // columns per operand per stage
  static constexpr int SFA_COL = 128;                       // [128, 128 + 2*SF_COLS)
  static constexpr int SFB_COL = 128 + 2 * 16;              // [160, 160 + 2*SF_COLS)
  static constexpr int TMEM_COLS = BLOCK_SCALED ? 256 : (N_MMA <= 32 ? 32 : (N_MMA <= 64 ? 64 : 128));
  static constexpr int MMA_PER_STAGE = 4;                   // K bytes per MMA = 32 (8-bit: 32 elems, fp4: 64)
};

struct Params {
  const float* x_scale;   // [M]     (rowwise kinds) or a_pts scalar (nvfp4) or null
  const float* w_scale;   // [N]     (rowwise kinds) or b_pts scalar (nvfp4) or null
  const __nv_bfloat16* bias;
  __nv_bfloat16* y;       // bf16 out [M, N]  (null when i32_out is set)
  int32_t* i32_out;       // raw int32 accumulators [M, N] (ao_int8_mm_i32)
  float* ws_partial;
  unsigned int* ws_sem;
  int M, N, K;            // K in ELEMENTS
  int splits;
  int sf_col_blocks_w;    // number of 4-wide scale column blocks per row block (blocked layout)
  int sf_col_blocks_x;
};

template <int KIND, int N_MMA>
__global__ void __launch_bounds__(NUM_THREADS)
lowp_linear_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_x,
                   const uint8_t* __restrict__ w_sf, const uint8_t* __restrict__ x_sf,
                   const Params p) {
  using C = Cfg<KIND, N_MMA>;
  constexpr int S = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * C::STAGE_BYTES);
  uint64_t* wfull = bars;            // weights (+ weight scales)
  uint64_t* xfull = bars + S;        // activations (+ activation scales)
  uint64_t* sempty = bars + 2 * S;   // MMA commit
  uint64_t* dfull = bars + 3 * S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dfull + 1);
  uint32_t* flag_slot = tmem_slot + 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x, split = blockIdx.y, m_blk = blockIdx.z;
  const int n0 = n_tile * ROWS, m0 = m_blk * N_MMA;
  constexpr int ELEMS_PER_BYTE = (KIND == KIND_NVF4) ? 2 : 1;
  const int k_bytes = p.K / ELEMS_PER_BYTE;
  const int total_chunks = k_bytes / KB;
  const int c_begin = (int)(((long long)total_chunks * split) / p.splits);
  const int c_end = (int)(((long long)total_chunks * (split + 1)) / p.splits);
  const int nchunks = c_end - c_begin;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&wfull[i], 1);
      mbar_init(&xfull[i], 1);
      mbar_init(&sempty[i], 1);
    }
    mbar_init(dfull, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_x);
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      const uint64_t pol_w = policy_evict_first();
      const uint64_t pol_x = policy_evict_last();
      auto issue_w = [&](int c) {
        const int s = c % S;
        uint8_t* st = smem + (size_t)s * C::STAGE_BYTES;
        const int kc = c_begin + c;
        mbar_expect_tx(&wfull[s], A_BYTES + C::SF_BYTES);
        tma_load_2d(st, &tm_w, &wfull[s], kc * (KB / (KIND == KIND_NVF4 ? 1 : 1)), n0, pol_w);
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
      auto issue_x = [&](int c) {
        const int s = c % S;
        uint8_t* st = smem + (size_t)s * C::STAGE_BYTES;
        const int kc = c_begin + c;
        mbar_expect_tx(&xfull[s], C::B_BYTES + C::SF_BYTES);
        tma_load_2d(st + A_BYTES, &tm_x, &xfull[s], kc * KB, m0, pol_x);
        if (C::BLOCK_SCALED) {
          const uint8_t* src = x_sf + ((size_t)((m0 / 128)) * p.sf_col_blocks_x + (size_t)kc * C::SF_TILES) * 512;
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                  smem_u32(st + A_BYTES + C::B_BYTES + C::SF_BYTES)),
              "l"(src), "r"(C::SF_BYTES), "r"(smem_u32(&xfull[s]))
              : "memory");
        }
      };
      const int pre = nchunks < S ? nchunks : S;
      for (int c = 0; c < pre; ++c) issue_w(c);
      pdl_wait();
      for (int c = 0; c < pre; ++c) issue_x(c);
      for (int c = S; c < nchunks; ++c) {
        mbar_wait(&sempty[c % S], ((c / S) & 1) ^ 1);
        issue_w(c);
        issue_x(c);
      }
    }
  } else if (warp == 1) {
    // instruction descriptor per kind
    constexpr uint32_t idesc =
        KIND == KIND_I8   ? make_idesc(2 /*s32*/, 1 /*int8*/, 1 /*int8*/, ROWS, N_MMA)
        : KIND == KIND_F8 ? make_idesc(1 /*f32*/, 0 /*e4m3*/, 0 /*e4m3*/, ROWS, N_MMA)
        : KIND == KIND_MXF8 ? make_idesc_bs(0 /*e4m3*/, 0 /*e4m3*/, 1 /*ue8m0*/, ROWS, N_MMA)
                            : make_idesc_bs(1 /*e2m1*/, 1 /*e2m1*/, 0 /*ue4m3*/, ROWS, N_MMA);
    for (int c = 0; c < nchunks; ++c) {
      const int s = c % S;
      mbar_wait(&wfull[s], (c / S) & 1);
      mbar_wait(&xfull[s], (c / S) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t ab = smem_u32(smem + (size_t)s * C::STAGE_BYTES);
        const uint32_t bb = ab + A_BYTES;
        uint32_t sfa_t = 0, sfb_t = 0;
        if (C::BLOCK_SCALED) {
          const int buf = c & 1;
          sfa_t = tmem_base + C::SFA_COL + buf * C::SF_COLS;
          sfb_t = tmem_base + C::SFB_COL + buf * C::SF_COLS;
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
          const uint32_t acc = (c > 0 || kk > 0) ? 1u : 0u;
          if (KIND == KIND_I8) mma_ss_i8(tmem_base, adesc, bdesc, idesc, acc);
          else if (KIND == KIND_F8) mma_ss_f8f6f4(tmem_base, adesc, bdesc, idesc, acc);
          else if (KIND == KIND_MXF8)
            // one scale byte per row per MMA: byte kk of the tile's column (sf_id fields)
            mma_ss_mxf8f6f4(tmem_base, adesc, bdesc, idesc | ((uint32_t)kk << 29) | ((uint32_t)kk << 4), acc, sfa_t, sfb_t);
          else
            // four scale bytes per row per MMA: tile kk
            mma_ss_mxf4nvf4_b16(tmem_base, adesc, bdesc, idesc, acc, sfa_t + kk * 4, sfb_t + kk * 4);
        }
        tc_commit(&sempty[s]);
        if (c == nchunks - 1) tc_commit(dfull);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;
    const int n = n0 + r;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    pdl_wait();
    mbar_wait(dfull, 0);
    tc_fence_after();
    const int tile_lin = m_blk * gridDim.x + n_tile;
    uint32_t* part = reinterpret_cast<uint32_t*>(p.ws_partial) + ((size_t)tile_lin * p.splits + split) * (N_MMA * ROWS);
    bool last = true;
    if (p.splits > 1) {
#pragma unroll
      for (int j = 0; j < N_MMA; j += 16) {
        uint32_t rr[16];
        tmem_ld_x16(lane_taddr + j, rr);
        tc_wait_ld();
#pragma unroll
        for (int q = 0; q < 16; ++q)
          if (m0 + j + q < p.M) __stcg(&part[(j + q) * ROWS + r], rr[q]);
      }
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) {
        const unsigned prev = atomicAdd(&p.ws_sem[tile_lin], 1u);
        *flag_slot = (prev == (unsigned)p.splits - 1) ? 1u : 0u;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      last = (*flag_slot != 0);
      if (last) __threadfence();
    }
    if (last && n < p.N) {
      const uint32_t* base = reinterpret_cast<const uint32_t*>(p.ws_partial) + (size_t)tile_lin * p.splits * (N_MMA * ROWS);
      float sw = 1.f, bias = 0.f;
      if (KIND == KIND_I8 || KIND == KIND_F8) sw = p.w_scale ? p.w_scale[n] : 1.f;
      if (KIND == KIND_NVF4) sw = (p.x_scale ? *p.x_scale : 1.f) * (p.w_scale ? *p.w_scale : 1.f);
      if (p.bias) bias = __bfloat162float(p.bias[n]);
#pragma unroll 1
      for (int j = 0; j < N_MMA; j += 16) {
        if (m0 + j >= p.M) break;
        uint32_t rr[16];
        float accf[16];
        int32_t acci[16];
        if (p.splits == 1) {
          tmem_ld_x16(lane_taddr + j, rr);
          tc_wait_ld();
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            acci[q] = (int32_t)rr[q];
            accf[q] = __uint_as_float(rr[q]);
          }
        } else {
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            acci[q] = 0;
            accf[q] = 0.f;
          }
          for (int sp = 0; sp < p.splits; ++sp) {  // fixed order: deterministic; 16 loads in flight
            const uint32_t* src = base + (size_t)sp * (N_MMA * ROWS) + (size_t)j * ROWS + r;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              const uint32_t v = __ldcg(src + q * ROWS);
              acci[q] += (int32_t)v;
              accf[q] += __uint_as_float(v);
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const int m = m0 + j + q;
          if (m >= p.M) continue;
          if (KIND == KIND_I8) {
            if (p.i32_out) {
              p.i32_out[(size_t)m * p.N + n] = acci[q];
              continue;
            }
            // int8/kernels.py:143-144 + int8_tensor.py:315-359: bf16 round between the scales
            const float t = __bfloat162float(__float2bfloat16_rn((float)acci[q] * p.x_scale[m]));
            p.y[(size_t)m * p.N + n] = __float2bfloat16_rn(t * sw + bias);
          } else if (KIND == KIND_F8) {
            p.y[(size_t)m * p.N + n] = __float2bfloat16_rn(accf[q] * (p.x_scale[m] * sw) + bias);
          } else if (KIND == KIND_MXF8) {
            p.y[(size_t)m * p.N + n] = __float2bfloat16_rn(accf[q] + bias);
          } else {
            p.y[(size_t)m * p.N + n] = __float2bfloat16_rn(accf[q] * sw + bias);
          }
        }
      }
    }
    if (p.splits > 1 && last) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) p.ws_sem[tile_lin] = 0;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
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
  const int n_tiles = ceil_div(N, ROWS);
  const int m_blocks = ceil_div(M, N_MMA);
  const int total_chunks = k_bytes / KB;
  int splits = (2 * sm_count()) / (n_tiles * m_blocks);
  if (splits < 1) splits = 1;
  if (splits > MAX_SPLITS) splits = MAX_SPLITS;
  if (splits > total_chunks) splits = total_chunks;
  // keep at least 4 chunks per split so the pipeline has something to stream
  while (splits > 1 && total_chunks / splits < 4) --splits;
  Params p;
  p.x_scale = x_scale; p.w_scale = w_scale;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.i32_out = i32_out;
  p.ws_sem = reinterpret_cast<unsigned int*>(ws);
  p.ws_partial = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + 64 * 1024);
  p.M = M; p.N = N; p.K = K; p.splits = splits;
  const int sf_per_row = KIND == KIND_MXF8 ? K / 32 : (KIND == KIND_NVF4 ? K / 16 : 0);
  p.sf_col_blocks_w = ceil_div(sf_per_row, 4);
  p.sf_col_blocks_x = ceil_div(sf_per_row, 4);
  if (splits > 1) {
    const size_t need = 64 * 1024 + (size_t)n_tiles * m_blocks * splits * N_MMA * ROWS * 4;
    if (!ws || ws_bytes < need || (size_t)n_tiles * m_blocks * 4 > 64 * 1024)
      return fail(AO_ERR_WORKSPACE, "lowp linear: workspace too small (%zu < %zu)", ws_bytes, need);
  }
  auto kern = lowp_linear_kernel<KIND, N_MMA>;
  static bool attr_set = false;
  if (!attr_set) {
    AO_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
    attr_set = true;
  }
  AO_CUDA_CHECK(ao::launch(kern, dim3(n_tiles, splits, m_blocks), dim3(NUM_THREADS), C::SMEM_BYTES, stream,
                           pdl_enabled(), tm_w, tm_x, w_sf, x_sf, p));
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
  AO_REQUIRE(K % 128 == 0, "int8 linear: K=%d must be a multiple of 128", K);
  AO_REQUIRE(N % 8 == 0, "int8 linear: N=%d must be a multiple of 8 (reference: int8/kernels.py:48-58)", N);
  if (M == 0) return AO_OK;
  AO_REQUIRE(xq && x_scale && wq && w_scale && y, "int8 linear: null pointer");
  return lowp::dispatch<lowp::KIND_I8>(reinterpret_cast<const uint8_t*>(xq), nullptr, x_scale, M, K,
                                       reinterpret_cast<const uint8_t*>(wq), nullptr, w_scale, N, bias,
                                       y, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int ao_int8_mm_i32(const int8_t* xq, int M, int K, const int8_t* wq, int N, int32_t* acc,
                              void* workspace, size_t workspace_bytes, void* stream) {
  AO_REQUIRE(M >= 0 && K > 0 && N > 0 && K % 128 == 0, "int8 mm: bad sizes M=%d K=%d N=%d", M, K, N);
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
  AO_REQUIRE(K % 128 == 0, "fp8 linear: K=%d must be a multiple of 128", K);
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
  AO_REQUIRE(K % 128 == 0, "mxfp8 linear: K=%d must be a multiple of 128", K);
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
