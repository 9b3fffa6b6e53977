// Can several threads issue tcgen05.mma concurrently?  NI issuer warps, each issues `iters` TS MMAs (M128,N16,K16)
// into its own accumulator; reports cycles per MMA per issuer and aggregate.
#include <cstdio>
#include <cuda_runtime.h>
#include "../ao_b200/csrc/ptx.cuh"
using namespace ao;
template <int NI, int NN>
__global__ void __launch_bounds__(256) bench(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[8];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  constexpr uint32_t idesc = make_idesc(1, 1, 1, 128, NN);
  if (warp < NI) {
    long long t0 = clock64();
    if (lane == 0) {
      const uint32_t b_s = smem_u32(smem);
      for (int i = 0; i < iters; ++i) {
        const int chunk = i >> 3, kk = i & 7;
        const uint32_t a_t = tmem + 256 + (chunk % 4) * 64 + kk * 8;
        const uint64_t bd = umma_desc_k_sw128(b_s + (kk >> 2) * (NN * 128) + (kk & 3) * 32);
        mma_ts_f16(tmem + warp * NN, a_t, bd, idesc, 1);
      }
      tc_commit(&bar[warp]);
    }
    __syncwarp();
    mbar_wait(&bar[warp], 0);
    if (lane == 0) out[blockIdx.x * 8 + warp] = clock64() - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}
template <int NI, int NN> void run(long long* d_out) {
  auto k = bench<NI, NN>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  const int iters = 2048;
  for (int rep = 0; rep < 2; ++rep) k<<<148, 256, 80 * 1024>>>(d_out, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148 * 8]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  double mx = 0; for (int b = 0; b < 148; ++b) for (int w = 0; w < NI; ++w) mx += h[b * 8 + w]; mx /= (148.0 * NI);
  printf("issuers=%d N=%3d: %7.1f cycles per MMA per issuer, %7.1f aggregate cycles/MMA (%s)\n", NI, NN, mx / iters, mx / iters / NI, cudaGetErrorString(e));
}
int main() {
  long long* d_out; cudaMalloc(&d_out, 148 * 8 * sizeof(long long));
  run<1, 16>(d_out); run<2, 16>(d_out); run<4, 16>(d_out);
  run<1, 32>(d_out); run<2, 32>(d_out); run<4, 32>(d_out);
  run<1, 128>(d_out); run<2, 128>(d_out);
  return 0;
}
