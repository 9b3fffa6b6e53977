// Is the ~70 cycles/MMA of a single issuing thread an ISSUE limit or the latency of the accumulate dependency
// chain?  One issuer rotates over NACC independent accumulators.  Also: cost of waits between 8-MMA batches.
//   WAIT: 0 none, 1 commit per batch, 2 commit + try_wait on an already-complete barrier, 3 commit + bar.sync
//   with a partner warp, 4 commit + try_wait on the PREVIOUS batch's commit barrier (real dependency, depth 2)
#include <cstdio>
#include <cuda_runtime.h>
#include "../ao_b200/csrc/ptx.cuh"
using namespace ao;
template <int NI, int NACC, bool SHARED, int NN, int WAIT>
__global__ void __launch_bounds__(256) bench(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[16];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x == 0) mbar_arrive(&bar[15]);   // phase 0 of bar[15] is complete for good
  __syncthreads();
  const uint32_t tmem = slot;
  constexpr uint32_t idesc = make_idesc(1, 1, 1, 128, NN);
  if (warp < NI) {
    long long t0 = clock64();
    const uint32_t b_s = smem_u32(smem);
    const uint32_t d0 = tmem + (SHARED ? 0 : warp * NACC * NN);
    for (int batch = 0; batch < iters / 8; ++batch) {
      if (lane == 0) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t a_t = tmem + 256 + (batch % 4) * 64 + kk * 8;
          const uint64_t bd = umma_desc_k_sw128(b_s + (kk >> 2) * (NN * 128) + (kk & 3) * 32);
          mma_ts_f16(d0 + (kk % NACC) * NN, a_t, bd, idesc, 1);
        }
        if (WAIT >= 1) tc_commit(&bar[8 + (batch & 1)]);
      }
      if (WAIT == 2) { while (!mbar_try_wait(&bar[15], 0)) {} }
      if (WAIT == 3) { asm volatile("bar.sync 1, 64;" ::: "memory"); }
      if (WAIT == 4 && batch >= 1) { while (!mbar_try_wait(&bar[8 + ((batch - 1) & 1)], ((batch - 1) >> 1) & 1)) {} }
      __syncwarp();
    }
    if (lane == 0) tc_commit(&bar[warp]);
    __syncwarp();
    mbar_wait(&bar[warp], 0);
    if (lane == 0) out[blockIdx.x * 8 + warp] = clock64() - t0;
  } else if (WAIT == 3 && warp == 4) {
    for (int batch = 0; batch < iters / 8; ++batch) asm volatile("bar.sync 1, 64;" ::: "memory");
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}
template <int NI, int NACC, bool SHARED, int NN, int WAIT> void run(long long* d_out) {
  auto k = bench<NI, NACC, SHARED, NN, WAIT>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  const int iters = 2048;
  for (int rep = 0; rep < 2; ++rep) k<<<148, 256, 80 * 1024>>>(d_out, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148 * 8]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  double mx = 0; for (int b = 0; b < 148; ++b) for (int w = 0; w < NI; ++w) mx += h[b * 8 + w]; mx /= (148.0 * NI);
  printf("issuers=%d nacc=%d shared=%d N=%3d wait=%d: %7.1f cycles/MMA per issuer, %7.1f aggregate, %7.0f per 8-MMA batch (%s)\n", NI, NACC, (int)SHARED, NN, WAIT,
         mx / iters, mx / iters / NI, mx / iters * 8 / NI, cudaGetErrorString(e));
}
int main() {
  long long* d_out; cudaMalloc(&d_out, 148 * 8 * sizeof(long long));
  run<1, 1, false, 16, 0>(d_out); run<1, 2, false, 16, 0>(d_out); run<1, 4, false, 16, 0>(d_out); run<1, 8, false, 16, 0>(d_out);
  run<1, 1, false, 32, 0>(d_out); run<1, 2, false, 32, 0>(d_out); run<1, 4, false, 32, 0>(d_out);
  run<1, 2, false, 64, 0>(d_out); run<1, 2, false, 128, 0>(d_out);
  run<3, 1, true, 16, 0>(d_out); run<3, 1, false, 16, 0>(d_out); run<3, 2, false, 16, 0>(d_out);
  printf("-- waits, 1 issuer, N=16\n");
  run<1, 1, false, 16, 1>(d_out); run<1, 1, false, 16, 2>(d_out); run<1, 1, false, 16, 3>(d_out); run<1, 1, false, 16, 4>(d_out);
  run<1, 4, false, 16, 1>(d_out); run<1, 4, false, 16, 2>(d_out); run<1, 4, false, 16, 3>(d_out); run<1, 4, false, 16, 4>(d_out);
  run<1, 2, false, 32, 2>(d_out); run<1, 2, false, 32, 4>(d_out);
  return 0;
}
