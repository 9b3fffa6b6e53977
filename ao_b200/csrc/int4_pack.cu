// tinygemm "tile_packed_to_4d" packing for int4 weights (setup-time kernels, bit-exact).
// Replaces aten._convert_weight_to_int4pack (PyTorch-core op the reference calls at
// torchao/quantization/quantize_/workflows/int4/int4_tile_packed_to_4d_tensor.py:198-204);
// layout documented at the top of int4_linear.cu.
#include <cuda_bf16.h>

#include "common.h"

namespace ao {

__device__ __forceinline__ uint32_t nib(const uint8_t* __restrict__ q, int K, int n, int k) {
  const uint8_t b = q[(size_t)n * (K / 2) + (k >> 1)];
  return (k & 1) ? (b & 15u) : (b >> 4);  // even k in the HIGH nibble of the pre-pack byte
}

__global__ void int4_pack_kernel(const uint8_t* __restrict__ q, int32_t* __restrict__ out, int N,
                                 int K, int ikt) {
  const int wpl = ikt / 2;  // words per lane
  const size_t total = (size_t)(N / 8) * (K / (ikt * 16)) * 32 * wpl;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int wd = idx % wpl;
  const int t = (idx / wpl) % 32;
  const size_t rest = idx / (wpl * 32);
  const int KT = K / (ikt * 16);
  const int ko = rest % KT;
  const int n8 = rest / KT;
  const int n = n8 * 8 + t / 4;
  const int k0 = (ikt * ko + 2 * wd) * 16 + 2 * (t % 4);
  uint32_t w = 0;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    w |= nib(q, K, n, k0 + 8 * e) << (4 * e);
    w |= nib(q, K, n, k0 + 8 * e + 1) << (16 + 4 * e);
  }
  out[idx] = (int32_t)w;
}

__global__ void int4_unpack_kernel(const int32_t* __restrict__ qd, uint8_t* __restrict__ q, int N,
                                   int K, int ikt) {
  // one thread per output byte (two nibbles k, k+1 with k even)
  const size_t total = (size_t)N * (K / 2);
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int n = idx / (K / 2);
  const int k = 2 * (idx % (K / 2));
  const int wpl = ikt / 2;
  const int KT = K / (ikt * 16);
  const int ko = k / (ikt * 16);
  const int kin = k % (ikt * 16);
  const int wd = kin / 32;
  const int k32 = kin % 32;
  const int e = k32 / 8;
  const int tq = (k32 % 8) / 2;
  const int t = (n % 8) * 4 + tq;
  const uint32_t w = (uint32_t)qd[(((size_t)(n / 8) * KT + ko) * 32 + t) * wpl + wd];
  const uint32_t lo = (w >> (4 * e)) & 15u;        // k
  const uint32_t hi = (w >> (16 + 4 * e)) & 15u;   // k + 1
  q[idx] = (uint8_t)((lo << 4) | hi);
}

__global__ void int4_dequant_kernel(const int32_t* __restrict__ qd,
                                    const __nv_bfloat16* __restrict__ sz,
                                    __nv_bfloat16* __restrict__ w_out, int N, int K, int g) {
  const size_t total = (size_t)N * (K / 2);
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int n = idx / (K / 2);
  const int k = 2 * (idx % (K / 2));
  const int KT = K / 128;
  const int ko = k / 128, kin = k % 128;
  const int wd = kin / 32, k32 = kin % 32, e = k32 / 8, tq = (k32 % 8) / 2;
  const int t = (n % 8) * 4 + tq;
  const uint32_t w = (uint32_t)qd[(((size_t)(n / 8) * KT + ko) * 32 + t) * 4 + wd];
  const int q0 = (w >> (4 * e)) & 15, q1 = (w >> (16 + 4 * e)) & 15;
  const size_t gi = ((size_t)(k / g) * N + n) * 2;
  const __nv_bfloat16 s = sz[gi], z = sz[gi + 1];
  w_out[(size_t)n * K + k] = __hfma(__int2bfloat16_rn(q0 - 8), s, z);
  w_out[(size_t)n * K + k + 1] = __hfma(__int2bfloat16_rn(q1 - 8), s, z);
}

}  // namespace ao

extern "C" {

int ao_int4_pack_tile4d(const uint8_t* q_u8, int32_t* qdata, int N, int K, int inner_k_tiles,
                        void* stream) {
  using namespace ao;
  AO_REQUIRE(inner_k_tiles == 2 || inner_k_tiles == 4 || inner_k_tiles == 8,
             "int4 pack: inner_k_tiles=%d not in {2,4,8}", inner_k_tiles);
  AO_REQUIRE(N > 0 && N % 8 == 0, "int4 pack: N=%d must be a positive multiple of 8", N);
  AO_REQUIRE(K > 0 && K % (inner_k_tiles * 16) == 0, "int4 pack: K=%d must be a multiple of %d", K,
             inner_k_tiles * 16);
  AO_REQUIRE(q_u8 && qdata, "int4 pack: null pointer");
  const size_t total = (size_t)N * K / 8;
  AO_CUDA_CHECK(launch(int4_pack_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0,
                       reinterpret_cast<cudaStream_t>(stream), false, q_u8, qdata, N, K,
                       inner_k_tiles));
  return AO_OK;
}

int ao_int4_unpack_tile4d(const int32_t* qdata, uint8_t* q_u8, int N, int K, int inner_k_tiles,
                          void* stream) {
  using namespace ao;
  AO_REQUIRE(inner_k_tiles == 2 || inner_k_tiles == 4 || inner_k_tiles == 8,
             "int4 unpack: inner_k_tiles=%d not in {2,4,8}", inner_k_tiles);
  AO_REQUIRE(N > 0 && N % 8 == 0 && K > 0 && K % (inner_k_tiles * 16) == 0,
             "int4 unpack: bad shape N=%d K=%d", N, K);
  AO_REQUIRE(q_u8 && qdata, "int4 unpack: null pointer");
  const size_t total = (size_t)N * K / 2;
  AO_CUDA_CHECK(launch(int4_unpack_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0,
                       reinterpret_cast<cudaStream_t>(stream), false, qdata, q_u8, N, K,
                       inner_k_tiles));
  return AO_OK;
}

int ao_int4_dequant_tile4d(const int32_t* qdata, const uint16_t* scale_and_zero, uint16_t* w_bf16,
                           int N, int K, int group_size, void* stream) {
  using namespace ao;
  AO_REQUIRE(N > 0 && N % 8 == 0 && K > 0 && K % 128 == 0, "int4 dequant: bad shape N=%d K=%d", N,
             K);
  AO_REQUIRE(group_size > 0 && K % group_size == 0 && group_size % 2 == 0,
             "int4 dequant: bad group_size=%d", group_size);
  AO_REQUIRE(qdata && scale_and_zero && w_bf16, "int4 dequant: null pointer");
  const size_t total = (size_t)N * K / 2;
  AO_CUDA_CHECK(launch(int4_dequant_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0,
                       reinterpret_cast<cudaStream_t>(stream), false, qdata,
                       reinterpret_cast<const __nv_bfloat16*>(scale_and_zero),
                       reinterpret_cast<__nv_bfloat16*>(w_bf16), N, K, group_size));
  return AO_OK;
}

}  // extern "C"
