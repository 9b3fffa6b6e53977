// How fast can ONE CTA per SM pull the int4 weight stream out of HBM, and with which copy engine?
//
// The access pattern is exactly that of ts_gemm.cuh on a 28672 x 4096 (gate|up) int4 g=32 weight: a chunk =
// 16 segments of 512 B (one per n8 row group, KT*512 B apart) + 4 segments of 512 B of (scale, zero) pairs
// (N*4 B apart); a CTA walks its stream-K range of (tile, kc) units.  The consumer does nothing: a stage is
// released as soon as it has landed, so the number printed is what the copy path can deliver with S stages
// (S x 10 KB) in flight per SM.
//
//   mode 0  TMA tensor maps (3-D box {32 words, 4, 16} + 2-D box {128 words, 4}), one mbarrier per stage
//   mode 2  cp.async (LDGSTS, 16 B per lane: one warp instruction = one 512 B segment), cp.async.mbarrier.arrive
//   mode 3  like 0 but TWO CTAs per SM (grid 296)
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o stream_microbench stream_microbench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int STAGE = 10 * 1024;
constexpr int MAXS = 20;

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(su32(b)) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(su32(b)), "r"(par) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) { while (!mbar_try(b, par)) {} }

struct P {
  const uint8_t* w;     // [N/8][KT][512]
  const uint8_t* sz;    // [K/32][N][4]
  int N, KT, n_tiles, stages, mode, hold;
  int producers, groups, hint, backoff, hold_b;   // producer warps (1-2), consumer groups of 4 warps (1-3), evict_first hint, nanosleep in waits
  unsigned long long* sink;
};

__global__ void __launch_bounds__(512) stream_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_sz, const __grid_constant__ CUtensorMap tm_w2, const P p) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + (size_t)p.stages * STAGE);
  uint64_t* empty = full + MAXS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages;
  const long long U = (long long)p.n_tiles * p.KT;
  const int G = gridDim.x, b = blockIdx.x;
  const int u0 = (int)(U * b / G), u1 = (int)(U * (b + 1) / G);
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], p.mode == 2 ? 32 : 1);
      mbar_init(&empty[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  auto wait = [&](uint64_t* bar, uint32_t par) {
    if (p.backoff) { while (!mbar_try(bar, par)) __nanosleep(p.backoff); }
    else mbar_wait(bar, par);
  };
  if (warp >= 12 && warp < 12 + p.producers) {
    // ---------------- producers: chunk i belongs to producer i % producers
    const int pi = warp - 12;
    for (int u = u0 + pi; u < u1; u += p.producers) {
      const int i = u - u0, s = i % S;
      if (i >= S) mbar_wait(&empty[s], ((i / S) & 1) ^ 1);
      const int tile = u / p.KT, kc = u % p.KT;
      uint8_t* st = smem + (size_t)s * STAGE;
      if (p.mode == 4) {
        if (lane == 0) {   // weight box = 16 rows of 512 B (2-D, no swizzle) instead of 64 rows of 128 B
          mbar_expect(&full[s], STAGE);
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(su32(st)),
                       "l"(&tm_w2), "r"(su32(&full[s])), "r"(kc * 128), "r"(tile * 16), "l"(pol) : "memory");
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(su32(st + 8192)),
                       "l"(&tm_sz), "r"(su32(&full[s])), "r"(tile * 128), "r"(kc * 4), "l"(pol) : "memory");
        }
      } else if (p.mode == 0 || p.mode == 3) {
        if (lane == 0) {
          mbar_expect(&full[s], STAGE);
          if (p.hint) {
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(su32(st)),
                         "l"(&tm_w), "r"(su32(&full[s])), "r"(0), "r"(4 * kc), "r"(tile * 16), "l"(pol) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(su32(st + 8192)),
                         "l"(&tm_sz), "r"(su32(&full[s])), "r"(tile * 128), "r"(kc * 4), "l"(pol) : "memory");
          } else {
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(su32(st)),
                         "l"(&tm_w), "r"(su32(&full[s])), "r"(0), "r"(4 * kc), "r"(tile * 16) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(su32(st + 8192)),
                         "l"(&tm_sz), "r"(su32(&full[s])), "r"(tile * 128), "r"(kc * 4) : "memory");
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 20; ++j) {
          const uint8_t* src = j < 16 ? p.w + (((size_t)(tile * 16 + j) * p.KT + kc) << 9)
                                      : p.sz + ((size_t)(kc * 4 + j - 16) * p.N + tile * 128) * 4;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(su32(st + j * 512 + lane * 16)), "l"(src + lane * 16) : "memory");
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(su32(&full[s])) : "memory");
      }
      __syncwarp();
    }
  } else if (warp < 4 * p.groups) {
    // ---------------- consumer groups of 4 warps (like the dequant warpgroups): chunk i belongs to group i % groups; all
    // four warps wait for the stage, hold it for p.hold cycles ("work"), warp 0 of the group releases it
    const int g = warp >> 2;
    unsigned long long acc = 0;
    for (int u = u0 + g; u < u1; u += p.groups) {
      const int i = u - u0, s = i % S;
      wait(&full[s], (i / S) & 1);
      acc += *(volatile uint32_t*)(smem + (size_t)s * STAGE + lane * 4);
      if (p.hold) {
        const long long t0 = clock64();
        while (clock64() - t0 < p.hold) {}
      }
      __syncwarp();
      if ((warp & 3) == 0 && lane == 0) mbar_arrive(&empty[s]);
      if (p.hold_b) {   // work that continues after the stage has been handed back (second half of a dequant)
        const long long t0 = clock64();
        while (clock64() - t0 < p.hold_b) {}
      }
    }
    if (acc == 0x1234567) p.sink[0] = acc;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int N = 114688, K = 4096, KT = K / 128, n_tiles = N / 128;
  const size_t wbytes = (size_t)N * K / 2, szbytes = (size_t)(K / 32) * N * 4;
  const int copies = 4;    // distinct weights per timing loop: >> L2
  uint8_t *w, *sz;
  CK(cudaMalloc(&w, wbytes * copies));
  CK(cudaMalloc(&sz, szbytes * copies));
  CK(cudaMemset(w, 1, wbytes * copies));
  CK(cudaMemset(sz, 1, szbytes * copies));
  unsigned long long* sink;
  CK(cudaMalloc(&sink, 8));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fn;
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  struct Cfg { int mode, S, prod, groups, hold, hint, backoff, hold_b; };
  // stages are a multiple of the consumer groups (a stage always belongs to the same group: parity waits cannot alias)
  const Cfg cfgs[] = {
      {0, 6, 1, 3, 0, 1, 0, 0},     {4, 6, 1, 3, 0, 1, 0, 0},     {0, 6, 2, 3, 0, 1, 0, 0},     {4, 6, 2, 3, 0, 1, 0, 0},
      {0, 6, 1, 3, 400, 1, 0, 600}, {4, 6, 1, 3, 400, 1, 0, 600}, {0, 6, 2, 3, 400, 1, 0, 600}, {4, 6, 2, 3, 400, 1, 0, 600},
      {4, 12, 1, 3, 400, 1, 0, 600}, {4, 12, 2, 3, 0, 1, 0, 0},
  };
  for (const Cfg& c : cfgs) {
    CUtensorMap tms[copies][3];
    for (int cc = 0; cc < copies; ++cc) {
      cuuint64_t d3[3] = {32, (cuuint64_t)4 * KT, (cuuint64_t)N / 8}, s3[2] = {128, (cuuint64_t)KT * 512};
      cuuint32_t b3[3] = {32, 4, 16}, e3[3] = {1, 1, 1};
      if (enc(&tms[cc][0], CU_TENSOR_MAP_DATA_TYPE_INT32, 3, w + wbytes * cc, d3, s3, b3, e3, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("enc w failed\n"); return 1; }
      cuuint64_t d2[2] = {(cuuint64_t)N, (cuuint64_t)K / 32}, s2[1] = {(cuuint64_t)N * 4};
      cuuint32_t b2[2] = {128, 4}, e2[2] = {1, 1};
      if (enc(&tms[cc][1], CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, sz + szbytes * cc, d2, s2, b2, e2, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("enc sz failed\n"); return 1; }
    }
    for (int cc = 0; cc < copies; ++cc) {
      cuuint64_t d2[2] = {(cuuint64_t)KT * 128, (cuuint64_t)N / 8}, s2[1] = {(cuuint64_t)KT * 512};
      cuuint32_t b2[2] = {128, 16}, e2[2] = {1, 1};
      if (enc(&tms[cc][2], CU_TENSOR_MAP_DATA_TYPE_INT32, 2, w + wbytes * cc, d2, s2, b2, e2, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("enc w2 failed\n"); return 1; }
    }
    const int grid = c.mode == 3 ? 2 * sms : sms;
    const size_t smem = (size_t)c.S * STAGE + 2 * MAXS * 8 + 1024 + 64;
    auto run = [&]() {
      for (int cc = 0; cc < copies; ++cc) {
        P p{w + wbytes * cc, sz + szbytes * cc, N, KT, n_tiles, c.S, c.mode, c.hold, c.prod, c.groups, c.hint, c.backoff, c.hold_b, sink};
        stream_kernel<<<grid, 512, smem>>>(tms[cc][0], tms[cc][1], tms[cc][2], p);
      }
    };
    run();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    run();
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double us = ms * 1e3 / copies, gb = (wbytes + szbytes) / 1e9;
    printf("mode %d stages %2d producers %d groups %d hold %4d+%4d hint %d backoff %2d grid %3d: %7.2f us/launch  %7.1f GB/s\n", c.mode, c.S, c.prod,
           c.groups, c.hold, c.hold_b, c.hint, c.backoff, grid, us, gb / (us * 1e-6));
    fflush(stdout);
  }
  return 0;
}
