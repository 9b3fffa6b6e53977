// tcgen05.mma issue-cost microbenchmark (sm_100a): cycles per MMA for the shapes the decode
// kernels use.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_microbench mma_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../ao_b200/csrc/ptx.cuh"
using namespace ao;

// MODE 0: TS f16 (A in TMEM)  1: SS f16  2: SS i8  3: SS f8f6f4
template <int MODE, int M, int N, int NACC>
__global__ void __launch_bounds__(128) bench(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0) {
    long long t0 = 0, t1 = 0;
    if (lane == 0) {
      constexpr uint32_t idesc = MODE <= 1 ? make_idesc(1, 1, 1, M, N) : (MODE == 2 ? make_idesc(2, 1, 1, M, N) : make_idesc(1, 0, 0, M, N));
      const uint32_t a_s = smem_u32(smem), b_s = smem_u32(smem + 32 * 1024);
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const uint32_t d = tmem + (i % NACC) * N;
        const uint64_t bd = umma_desc_k_sw128(b_s + (i & 3) * 32);
        if (MODE == 0) mma_ts_f16(d, tmem + 256 + (i & 7) * 8, bd, idesc, 1);
        else {
          const uint64_t ad = umma_desc_k_sw128(a_s + (i & 3) * 32);
          if (MODE == 1) mma_ss_f16(d, ad, bd, idesc, 1);
          else if (MODE == 2) mma_ss_i8(d, ad, bd, idesc, 1);
          else mma_ss_f8f6f4(d, ad, bd, idesc, 1);
        }
      }
      tc_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    if (lane == 0) { t1 = clock64(); out[blockIdx.x] = t1 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

template <int MODE, int M, int N, int NACC>
void run(const char* name, long long* d_out) {
  auto k = bench<MODE, M, N, NACC>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 2048;
  for (int rep = 0; rep < 2; ++rep) k<<<148, 128, 64 * 1024>>>(d_out, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  printf("%-28s M=%3d N=%3d nacc=%d : %7.1f cycles/MMA  (%s)\n", name, M, N, NACC, avg / iters, cudaGetErrorString(e));
}

int main() {
  long long* d_out; cudaMalloc(&d_out, 148 * sizeof(long long));
  run<0, 128, 16, 1>("TS f16", d_out);
  run<0, 128, 16, 4>("TS f16", d_out);
  run<0, 128, 32, 1>("TS f16", d_out);
  run<0, 128, 64, 1>("TS f16", d_out);
  run<0, 128, 128, 1>("TS f16", d_out);
  run<0, 64, 16, 1>("TS f16", d_out);
  run<0, 64, 32, 1>("TS f16", d_out);
  run<1, 128, 16, 1>("SS f16", d_out);
  run<1, 128, 32, 1>("SS f16", d_out);
  run<1, 128, 128, 1>("SS f16", d_out);
  run<1, 128, 256, 1>("SS f16", d_out);
  run<1, 64, 16, 1>("SS f16", d_out);
  run<2, 128, 16, 1>("SS i8", d_out);
  run<2, 128, 32, 1>("SS i8", d_out);
  run<2, 128, 128, 1>("SS i8", d_out);
  run<3, 128, 16, 1>("SS f8f6f4", d_out);
  run<3, 128, 32, 1>("SS f8f6f4", d_out);
  run<3, 128, 128, 1>("SS f8f6f4", d_out);
  run<3, 64, 16, 1>("SS f8f6f4", d_out);
  return 0;
}
