#include <cstdio>
#include <cuda_runtime.h>
#include "../ao_b200/csrc/ptx.cuh"
using namespace ao;
__device__ __forceinline__ bool elect1() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
template <int MODE>
__global__ void __launch_bounds__(256) bench(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[8];
  __shared__ uint32_t slot;
  const int warp = MODE == 0 ? (threadIdx.x >> 5) : __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = MODE == 0 ? slot : __shfl_sync(0xffffffffu, slot, 0);
  constexpr uint32_t idesc = make_idesc(1, 1, 1, 128, 16);
  if (warp == 1) {
    long long t0 = clock64();
    const uint32_t b_s = smem_u32(smem);
    if (MODE == 0) {
      if (lane == 0) {
        for (int batch = 0; batch < iters / 8; ++batch) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint32_t a_t = tmem + 256 + (batch % 4) * 64 + kk * 8;
            const uint64_t bd = umma_desc_k_sw128(b_s + (kk >> 2) * (16 * 128) + (kk & 3) * 32);
            mma_ts_f16(tmem, a_t, bd, idesc, 1);
          }
          tc_commit(&bar[2 + (batch & 1)]);
        }
      }
    } else {
      for (int batch = 0; batch < iters / 8; ++batch) {
        if (elect1()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint32_t a_t = tmem + 256 + (batch % 4) * 64 + kk * 8;
            const uint64_t bd = umma_desc_k_sw128(b_s + (kk >> 2) * (16 * 128) + (kk & 3) * 32);
            mma_ts_f16(tmem, a_t, bd, idesc, 1);
          }
          tc_commit(&bar[2 + (batch & 1)]);
        }
        __syncwarp();
      }
    }
    if (lane == 0) tc_commit(&bar[0]);
    __syncwarp();
    mbar_wait(&bar[0], 0);
    if (lane == 0) out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}
template <int MODE> void run(long long* d_out) {
  auto k = bench<MODE>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  const int iters = 2048;
  for (int rep = 0; rep < 2; ++rep) k<<<148, 256, 80 * 1024>>>(d_out, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  double mx = 0; for (int b = 0; b < 148; ++b) mx += h[b]; mx /= 148.0;
  printf("mode=%d: %7.1f cycles/MMA (%s)\n", MODE, mx / iters, cudaGetErrorString(e));
}
int main() {
  long long* d_out; cudaMalloc(&d_out, 148 * sizeof(long long));
  run<0>(d_out); run<1>(d_out);
  return 0;
}
