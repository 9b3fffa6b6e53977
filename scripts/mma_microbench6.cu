// What slows the MMA issuer inside the real kernel?  One issuer (8 MMAs + commit + try_wait per batch, N=16)
// while 8 background warps (two per scheduler... warps 4..11) hammer one resource:
//   BG 0 nothing, 1 tcgen05.st back-to-back, 2 dependent-free HFMA2 ALU loop, 3 ld.shared.v4 loop,
//   4 realistic mix: ~700 cycles of ALU then 64 columns of tcgen05.st, 5 mbarrier try_wait polling (never completes)
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../ao_b200/csrc/ptx.cuh"
using namespace ao;
template <int BG, int WAIT>
__global__ void __launch_bounds__(512) bench(long long* out, int iters, unsigned* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[16];
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 512) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); stop = 0; }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x == 0) mbar_arrive(&bar[15]);
  __syncthreads();
  const uint32_t tmem = slot;
  constexpr uint32_t idesc = make_idesc(1, 1, 1, 128, 16);
  if (warp == 13) {
    long long t0 = clock64();
    const uint32_t b_s = smem_u32(smem);
    for (int batch = 0; batch < iters / 8; ++batch) {
      if (lane == 0) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t a_t = tmem + 64 + (batch % 3) * 64 + kk * 8;
          const uint64_t bd = umma_desc_k_sw128(b_s + (kk >> 2) * (16 * 128) + (kk & 3) * 32);
          mma_ts_f16(tmem, a_t, bd, idesc, 1);
        }
        tc_commit(&bar[8 + (batch & 1)]);
      }
      if (WAIT == 2) { while (!mbar_try_wait(&bar[15], 0)) {} }
      if (WAIT == 4 && batch >= 1) { while (!mbar_try_wait(&bar[8 + ((batch - 1) & 1)], ((batch - 1) >> 1) & 1)) {} }
      __syncwarp();
    }
    if (lane == 0) tc_commit(&bar[0]);
    __syncwarp();
    mbar_wait(&bar[0], 0);
    if (lane == 0) { out[blockIdx.x] = clock64() - t0; stop = 1; }
  } else if (warp < 8 && BG != 0) {
    const uint32_t la = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256 + (warp >> 2) * 64;  // scratch columns 256..383
    uint32_t v[32];
    for (int i = 0; i < 32; ++i) v[i] = lane * 77 + i;
    __nv_bfloat162 h[8];
    for (int i = 0; i < 8; ++i) h[i] = __floats2bfloat162_rn(1.f + i, 2.f);
    unsigned acc = 0;
    while (!stop) {
      if (BG == 1) { tmem_st_x32(la, v); tmem_st_x32(la + 32, v); tc_wait_st(); }
      if (BG == 2 || BG == 4) {
        for (int r = 0; r < (BG == 4 ? 44 : 64); ++r)
#pragma unroll
          for (int i = 0; i < 8; ++i) h[i] = __hfma2(h[i], h[(i + 1) & 7], h[i]);
      }
      if (BG == 4) { tmem_st_x32(la, v); tmem_st_x32(la + 32, v); tc_wait_st(); }
      if (BG == 3) {
        for (int r = 0; r < 16; ++r) { uint4 q = *reinterpret_cast<uint4*>(smem + ((lane * 16 + r * 512 + warp * 8192) & 65535)); acc += q.x ^ q.y ^ q.z ^ q.w; }
      }
      if (BG == 5) { acc += mbar_try_wait(&bar[14], 0) ? 1 : 0; }
    }
    for (int i = 0; i < 8; ++i) acc += *reinterpret_cast<unsigned*>(&h[i]);
    if (acc == 0x12345678) sink[0] = acc;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}
template <int BG, int WAIT> void run(long long* d_out, unsigned* sink) {
  auto k = bench<BG, WAIT>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  const int iters = 1024;
  for (int rep = 0; rep < 2; ++rep) k<<<148, 512, 80 * 1024>>>(d_out, iters, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  double mx = 0; for (int b = 0; b < 148; ++b) mx += h[b]; mx /= 148.0;
  printf("bg=%d wait=%d: %7.1f cycles/MMA, %7.0f per 8-MMA batch (%s)\n", BG, WAIT, mx / iters, mx / iters * 8, cudaGetErrorString(e));
}
int main() {
  long long* d_out; cudaMalloc(&d_out, 148 * sizeof(long long));
  unsigned* sink; cudaMalloc(&sink, 4);
  run<0, 2>(d_out, sink); run<1, 2>(d_out, sink); run<2, 2>(d_out, sink); run<3, 2>(d_out, sink); run<4, 2>(d_out, sink); run<5, 2>(d_out, sink);
  run<0, 4>(d_out, sink); run<1, 4>(d_out, sink); run<2, 4>(d_out, sink); run<3, 4>(d_out, sink); run<4, 4>(d_out, sink); run<5, 4>(d_out, sink);
  return 0;
}
