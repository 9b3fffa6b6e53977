// Is it safe for several threads to issue tcgen05.mma that accumulate into the SAME TMEM accumulator?
// A (TMEM) = ones, B (smem) = ones, D zeroed; NI issuers x `iters` MMAs (K=16) => every D element must be 16*NI*iters.
#include <cstdio>
#include <cuda_runtime.h>
#include "../ao_b200/csrc/ptx.cuh"
using namespace ao;
template <int NI, int NN>
__global__ void __launch_bounds__(256) bench(int* bad, float* sample, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[8];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0x3f803f80u;  // bf16 ones
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (warp < 4) {  // fill A region (cols 256..511) with ones and zero D (cols 0..127)
    uint32_t ones[32], zeros[32];
    for (int i = 0; i < 32; ++i) { ones[i] = 0x3f803f80u; zeros[i] = 0; }
    const uint32_t la = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 256; c += 32) tmem_st_x32(la + 256 + c, ones);
    for (int c = 0; c < 128; c += 32) tmem_st_x32(la + c, zeros);
    tc_wait_st();
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  constexpr uint32_t idesc = make_idesc(1, 1, 1, 128, NN);
  if (warp < NI) {
    if (lane == 0) {
      const uint32_t b_s = smem_u32(smem);
      for (int i = 0; i < iters; ++i) {
        const int chunk = i >> 3, kk = i & 7;
        const uint32_t a_t = tmem + 256 + ((chunk + warp) % 4) * 64 + kk * 8;
        const uint64_t bd = umma_desc_k_sw128(b_s + (kk >> 2) * (NN * 128) + (kk & 3) * 32);
        mma_ts_f16(tmem, a_t, bd, idesc, 1);   // everyone accumulates into D at column 0
      }
      tc_commit(&bar[warp]);
    }
    __syncwarp();
    mbar_wait(&bar[warp], 0);
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp < 4) {
    uint32_t rr[16];
    tmem_ld_x16(tmem + ((uint32_t)(warp * 32) << 16), rr);
    tc_wait_ld();
    const float expect = 16.f * NI * iters;
    int nb = 0;
    for (int q = 0; q < 16 && q < NN; ++q) if (__uint_as_float(rr[q]) != expect) ++nb;
    if (nb) atomicAdd(bad, nb);
    if (blockIdx.x == 0 && threadIdx.x == 0) { sample[0] = __uint_as_float(rr[0]); sample[1] = expect; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}
template <int NI, int NN> void run(int* d_bad, float* d_s) {
  auto k = bench<NI, NN>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  cudaMemset(d_bad, 0, 4);
  k<<<148, 256, 80 * 1024>>>(d_bad, d_s, 256);
  cudaError_t e = cudaDeviceSynchronize();
  int bad; float s[2]; cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost); cudaMemcpy(s, d_s, 8, cudaMemcpyDeviceToHost);
  printf("issuers=%d N=%3d: wrong elements=%d  (sample D=%.0f expected %.0f) %s\n", NI, NN, bad, s[0], s[1], cudaGetErrorString(e));
}
int main() {
  int* d_bad; float* d_s; cudaMalloc(&d_bad, 4); cudaMalloc(&d_s, 8);
  run<1, 16>(d_bad, d_s); run<2, 16>(d_bad, d_s); run<4, 16>(d_bad, d_s); run<3, 32>(d_bad, d_s); run<4, 32>(d_bad, d_s); run<4, 64>(d_bad, d_s);
  return 0;
}
