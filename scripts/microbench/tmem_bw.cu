// TMEM -> register read bandwidth and latency on sm_100a (tcgen05.ld), 1 CTA on one SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I csrc scripts/microbench/tmem_bw.cu -o /tmp/tmem_bw && /tmp/tmem_bw
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace pb;

template <int COLS>   // columns per load instruction: 16 or 32
__global__ void __launch_bounds__(512, 1) k_ld(long long* out, int iters, int warps_active) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot + (uint32_t((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < warps_active) {
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int c = 0; c < 128; c += COLS) {
        if constexpr (COLS == 32) {
          uint32_t r[32];
          tmem_ld_32x32(tm + ((warp >> 2) & 1) * 128 + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 32; ++t) acc ^= r[t];
        } else {
          uint32_t r[16];
          tmem_ld_32x16(tm + ((warp >> 2) & 1) * 128 + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 16; ++t) acc ^= r[t];
        }
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345) out[1] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(slot);
}

// same but 4 loads in flight before the wait
__global__ void __launch_bounds__(512, 1) k_ld_deep(long long* out, int iters, int warps_active) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot + (uint32_t((warp & 3) * 32) << 16) + ((warp >> 2) & 1) * 128;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < warps_active) {
    for (int i = 0; i < iters; ++i) {
      uint32_t r0[32], r1[32], r2[32], r3[32];
      tmem_ld_32x32(tm, r0);
      tmem_ld_32x32(tm + 32, r1);
      tmem_ld_32x32(tm + 64, r2);
      tmem_ld_32x32(tm + 96, r3);
      tmem_ld_wait();
#pragma unroll
      for (int t = 0; t < 32; ++t) acc ^= r0[t] ^ r1[t] ^ r2[t] ^ r3[t];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345) out[1] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(slot);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  const int iters = 2000;
  for (int wa : {1, 4, 8, 16}) {
    long long h = 0;
    k_ld<32><<<1, 512>>>(d, iters, wa);
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    double bytes = double(wa) * iters * 128 * 32 * 4;
    printf("x32 ld+wait      warps=%2d  clk/iter(128 cols)=%7.1f  B/clk/SM=%7.1f\n", wa, double(h) / iters, bytes / h);
    k_ld<16><<<1, 512>>>(d, iters, wa);
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("x16 ld+wait      warps=%2d  clk/iter(128 cols)=%7.1f  B/clk/SM=%7.1f\n", wa, double(h) / iters, bytes / h);
    k_ld_deep<<<1, 512>>>(d, iters, wa);
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("4 x x32 in flight warps=%2d  clk/iter(128 cols)=%7.1f  B/clk/SM=%7.1f\n", wa, double(h) / iters, bytes / h);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
