// Follow-up microbenchmark: why do TS-mode MMAs cost ~230 cycles each inside the real kernel?
// Variants (all M=128, N=16, kind::f16, 2048 MMAs issued by one thread, timed issue->completion):
//  0 same 8 A tiles reused (like the first microbench)       1 A rotating over 6 stages x 64 columns
//  2 = 1 + commit (to a dummy mbarrier) after every 8 MMAs     3 = 2 + fence::after_thread_sync + mbarrier try_wait per 8
//  4 = 1 + 8 other warps hammering tcgen05.st into other TMEM columns
//  5 SS mode, A rotating over 4 smem tiles                     6 = 1 but B descriptor rotating over 8 smem tiles
#include <cstdio>
#include <cuda_runtime.h>
#include "../ao_b200/csrc/ptx.cuh"
using namespace ao;

__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

template <int V, int NN = 16>
__global__ void __launch_bounds__(320) bench(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, dummy[8];
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  __shared__ volatile int done_upto;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 320) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) mbar_init(&dummy[i], 1); stop = 0; done_upto = 0; fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  constexpr uint32_t idesc = make_idesc(1, 1, 1, 128, NN);
  if (warp == 0) {
    long long t0 = 0, t1 = 0;
    if (lane == 0) {
      const uint32_t a_s = smem_u32(smem), b_s = smem_u32(smem + (NN > 16 ? 0 : 128 * 1024));
      t0 = clock64();
      int ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int i = 0; i < iters; ++i) {
        const int chunk = i >> 3, kk = i & 7;
        uint32_t a_t = (V == 0) ? tmem + 256 + kk * 8 : tmem + 256 + (chunk % 4) * 64 + kk * 8;
        uint64_t bd = umma_desc_k_sw128(b_s + (V == 6 ? (chunk & 7) * 4096 : 0) + (kk >> 2) * (NN * 128) + (kk & 3) * 32);
        const uint32_t d = tmem + (NN <= 128 ? (chunk & 1) * NN : 0);
        if (V == 5) mma_ss_f16(d, umma_desc_k_sw128(a_s + (chunk & 3) * 32768 + (kk >> 2) * 16384 + (kk & 3) * 32), bd, idesc, 1);
        else mma_ts_f16(d, a_t, bd, idesc, 1);
        if (V == 19 && kk == 7) { tc_commit(&dummy[chunk & 7]); if (blockIdx.x == 0 && chunk < 48) out[200 + chunk] = clock64() - t0; }
        if ((V == 17 || V == 18) && (i & 31) == 31) tc_commit(&dummy[(i >> 5) & 7]);
        if ((V == 2 || V == 3 || (V >= 7 && V != 17 && V != 18 && V != 19)) && kk == 7) {
          tc_commit(&dummy[chunk & 7]);
          if (V == 7) tc_fence_after();
          if (V == 8) { if (chunk >= 4) { const int c = (chunk - 4) & 7; mbar_wait(&dummy[c], ph[c]); ph[c] ^= 1; } }
          if (V == 9) { if (chunk >= 1) { const int c = (chunk - 1) & 7; mbar_wait(&dummy[c], ph[c]); ph[c] ^= 1; } }
          if (V == 11) { if (chunk >= 1) { const int c = (chunk - 1) & 7; while (!mbar_test_wait(&dummy[c], ph[c])) {} ph[c] ^= 1; } }
          if (V == 12) { const int c = chunk & 7; while (!mbar_test_wait(&dummy[c], ph[c])) {} ph[c] ^= 1; }
          if (V == 13) { if (chunk >= 4) { const int c = (chunk - 4) & 7; while (!mbar_test_wait(&dummy[c], ph[c])) {} ph[c] ^= 1; } }
          if (V == 20) { if ((chunk & 3) == 3) { while (done_upto < chunk - 3) {} } }   // poll once per 4 chunks
          if (V == 21) { if ((chunk & 3) == 3) { if (chunk >= 4) { const int c = (chunk - 4) & 7; mbar_wait(&dummy[c], ph[c]); ph[c] ^= 1; } } }
          if (V == 22) { asm volatile("nanosleep.u32 20;" ::: "memory"); }   // non-memory pause per chunk
          if (V == 23) { volatile int* q = &stop; int v = *q; if (v == 12345) out[400] = v; }  // one plain LDS per chunk, no dependence
          if (V == 14) { while (done_upto < chunk) {} }        // chunk-1 complete, observed through shared memory
          if (V == 15) { while (done_upto < chunk - 3) {} }    // chunk-4 complete
          if (V == 10) { const int c = chunk & 7; mbar_wait(&dummy[c], ph[c]); ph[c] ^= 1; }
          if (V == 3) {
            // wait for the commit of 4 chunks ago, like the pipeline's aempty/afull handshakes
            if (chunk >= 4) { const int c = (chunk - 4) & 7; mbar_wait(&dummy[c], ph[c]); ph[c] ^= 1; }
            tc_fence_after();
          }
        }
      }
      tc_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    if (lane == 0) { t1 = clock64(); out[blockIdx.x] = t1 - t0; stop = 1; }
  } else if (V == 19 && warp == 1) {
    if (lane == 0) {
      int ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const long long tb = clock64();
      for (int c = 0; c < 48; ++c) { mbar_wait(&dummy[c & 7], ph[c & 7]); ph[c & 7] ^= 1; if (blockIdx.x == 0) out[300 + c] = clock64() - tb; }
    }
  } else if (V == 17 && warp == 1) {
    if (lane == 0) {
      int ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int c = 0; c < iters / 32; ++c) { mbar_wait(&dummy[c & 7], ph[c & 7]); ph[c & 7] ^= 1; done_upto = c + 1; }
    }
  } else if ((V == 14 || V == 15 || V == 16 || V == 20) && warp == 1) {
    // sentinel: waits on the commit barriers in order and publishes progress through shared memory
    if (lane == 0) {
      int ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int c = 0; c < iters / 8; ++c) { mbar_wait(&dummy[c & 7], ph[c & 7]); ph[c & 7] ^= 1; done_upto = c + 1; }
    }
  } else if (V == 4 && warp >= 2) {
    // warps 2..9 (two warpgroups): keep writing 32 columns per thread into TMEM columns [0,...) not used as A
    uint32_t r[32];
    for (int i = 0; i < 32; ++i) r[i] = i * 0x3f803f80u;
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 32 + ((warp - 2) >> 2) * 32;
    while (!stop) { tmem_st_x32(taddr, r); tc_wait_st(); }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

template <int V, int NN = 16> void run(const char* name, long long* d_out) {
  auto k = bench<V, NN>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
  const int iters = 2048;
  for (int rep = 0; rep < 2; ++rep) k<<<148, 320, 170 * 1024>>>(d_out, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  printf("N=%3d V%d %-52s: %7.1f cycles/MMA (%s)\n", NN, V, name, avg / iters, cudaGetErrorString(e));
}
int main() {
  long long* d_out; cudaMalloc(&d_out, 512 * sizeof(long long));
  run<2, 128>("TS N=128 commit per 8 (issue rate)", d_out);
  run<14, 128>("TS N=128 sentinel, 1 back (true rate)", d_out);
  run<10, 128>("TS N=128 wait same chunk (latency)", d_out);
  run<2, 64>("TS N=64 commit per 8 (issue rate)", d_out);
  run<14, 64>("TS N=64 sentinel, 1 back (true rate)", d_out);
  run<14, 32>("TS N=32 sentinel, 1 back (true rate)", d_out);
  run<14, 256>("TS N=256 sentinel, 1 back (true rate)", d_out);
  run<16>("TS N=16 commit per 8 + sentinel waits, issuer never waits", d_out);
  run<17>("TS N=16 commit per 32 + sentinel waits, issuer never waits", d_out);
  run<18>("TS N=16 commit per 32, nobody waits", d_out);
  run<20>("TS N=16 commit per 8; issuer polls smem flag once per 4 chunks", d_out);
  run<21>("TS N=16 commit per 8; issuer mbar_wait once per 4 chunks (4 back)", d_out);
  run<22>("TS N=16 commit per 8; nanosleep 20 per chunk", d_out);
  run<23>("TS N=16 commit per 8; one independent LDS per chunk", d_out);
  run<19>("TS N=16 commit per 8, sentinel timestamps completions", d_out);
  { long long h[512]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("chunk: issue_done / commit_seen (cycles, both from ~kernel start of their warps)\n");
    for (int c = 0; c < 48; c += 1) printf("  %2d: %6lld / %6lld\n", c, h[200 + c], h[300 + c]); }
  run<0>("TS same 8 A tiles", d_out);
  run<1>("TS A rotating over 6 stages", d_out);
  run<2>("TS rotating + commit per 8", d_out);
  run<3>("TS rotating + commit + fence + wait(4 chunks back) per 8", d_out);
  run<4>("TS rotating + 8 warps tcgen05.st contention", d_out);
  run<5>("SS A rotating over 4 smem tiles", d_out);
  run<6>("TS rotating + B rotating over 8 smem tiles", d_out);
  run<7>("TS rotating + commit + fence::after per 8 (no wait)", d_out);
  run<8>("TS rotating + commit + wait(4 chunks back) per 8 (no fence)", d_out);
  run<9>("TS rotating + commit + wait(1 chunk back) per 8", d_out);
  run<10>("TS rotating + commit + wait(same chunk) per 8 [latency]", d_out);
  run<14>("TS rotating + commit; sentinel warp waits, issuer polls smem flag (1 back)", d_out);
  run<15>("TS rotating + commit; sentinel warp waits, issuer polls smem flag (4 back)", d_out);
  run<11>("TS rotating + commit + TEST_wait spin (1 chunk back)", d_out);
  run<12>("TS rotating + commit + TEST_wait spin (same chunk) [latency]", d_out);
  run<13>("TS rotating + commit + TEST_wait spin (4 chunks back)", d_out);
  return 0;
}
