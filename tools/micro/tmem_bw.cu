// Micro-benchmark: TMEM read bandwidth of tcgen05.ld.32x32b.x32 with 4 / 8 / 16 warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I sct_gan_b200/csrc tools/micro/tmem_bw.cu -o gpurun_out/tmem_bw
#include <cstdio>
#include "common.cuh"
using namespace sct;
__global__ void __launch_bounds__(512, 1) k(long long* out, float* sink, int iters, int ld_per_iter) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(smem_u32(&tptr), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tptr;
  const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int c = 0; c < ld_per_iter; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem + lane_sel + ((c + warp / 4) % 16) * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc += __uint_as_float(r[i]);
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
__global__ void __launch_bounds__(512, 1) k4(long long* out, float* sink, int iters) {  // 4 loads in flight per wait
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(smem_u32(&tptr), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tptr;
  const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t r0[32], r1[32], r2[32], r3[32];
    const uint32_t b = tmem + lane_sel + ((warp / 4) % 4) * 128;
    tmem_ld32(b, r0); tmem_ld32(b + 32, r1); tmem_ld32(b + 64, r2); tmem_ld32(b + 96, r3);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) acc += __uint_as_float(r0[i]) + __uint_as_float(r1[i]) + __uint_as_float(r2[i]) + __uint_as_float(r3[i]);
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
namespace sct { void set_error(const char*, ...) {} int num_sms() { return 148; } }
int main() {
  long long* d; float* s; cudaMalloc(&d, 148 * 8); cudaMalloc(&s, 4);
  const int iters = 2000;
  for (int threads : {128, 256, 512}) {
    k<<<148, threads>>>(d, s, iters, 1);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double bytes = (double)(threads / 32) * iters * 4096.0;
    printf("1 ld/wait, %2d warps/SM: %lld clks, %.1f B/clk/SM\n", threads / 32, h[0], bytes / h[0]);
    k4<<<148, threads>>>(d, s, iters);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    bytes = (double)(threads / 32) * iters * 4 * 4096.0;
    printf("4 ld/wait, %2d warps/SM: %lld clks, %.1f B/clk/SM   err=%s\n", threads / 32, h[0], bytes / h[0], cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
