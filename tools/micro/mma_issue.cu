// Micro-benchmark: cost of a tcgen05.mma (M = 128, K = 16, bf16) as a function of N, the operand sources and the
// number of issuing warps.  One CTA per SM; the operands are whatever bytes are in shared memory (timing only).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I sct_gan_b200/csrc tools/micro/mma_issue.cu -o gpurun_out/mma_issue
#include <cstdio>
#include "common.cuh"
using namespace sct;

// mode 0: A, B from shared memory (K-major, swizzle-128); mode 1: A from TMEM, B K-major swizzle-128;
// mode 2: A, B K-major swizzle-64 (32-column blocks, the attention operand tiles); mode 3: A from TMEM, B MN-major
// swizzle-64 (the dV / dK products of the attention backward)
template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters, int issuers) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t tptr;
  __shared__ uint64_t bars[4];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars[i]), 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tptr), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tptr;
  constexpr uint32_t idesc = umma_idesc_bf16(128, N, false, false);
  constexpr uint32_t kHi = umma_desc_hi(1024, UMMA_SW128);
  long long t0 = 0, t1 = 0;
  if (warp < issuers && lane == 0) {
    const uint32_t sA = base + warp * 65536, sB = sA + 16384;  // [128 x 64] A tile, [N x 64] B tile per issuer
    const uint32_t d = tmem + warp * 128;                       // own accumulator
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        constexpr uint32_t kHi64 = umma_desc_hi(512, UMMA_SW64);
        constexpr uint32_t idesc_mn = umma_idesc_bf16(128, N, false, true);
        if (MODE == 0)
          tc_mma_bf16_lh(d, umma_desc_lo(sA + kk * 32, 16), kHi, umma_desc_lo(sB + kk * 32, 16), kHi, idesc, 1u);
        else if (MODE == 1)
          tc_mma_bf16_ts(d, tmem + 256 + warp * 64 + kk * 8, umma_desc_lo(sB + kk * 32, 16), kHi, idesc, 1u);
        else if (MODE == 2)
          tc_mma_bf16_lh(d, umma_desc_lo(sA + (kk >> 1) * 8192 + (kk & 1) * 32, 16), kHi64,
                         umma_desc_lo(sB + (kk >> 1) * 8192 + (kk & 1) * 32, 16), kHi64, idesc, 1u);
        else
          tc_mma_bf16_ts(d, tmem + 256 + warp * 64 + kk * 8, umma_desc_lo(sB + kk * 1024, 8192), kHi64, idesc_mn, 1u);
      }
    }
    tc_commit(smem_u32(&bars[warp]));
    t1 = clock64();  // issue time
    mbar_wait(smem_u32(&bars[warp]), 0);
    const long long t2 = clock64();  // all retired
    if (blockIdx.x == 0) {
      out[2 * warp] = t1 - t0;
      out[2 * warp + 1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

namespace sct {
void set_error(const char*, ...) {}
int num_sms() { return 148; }
}  // namespace sct

template <int N, int MODE>
void run(long long* d, const char* what) {
  const int iters = 500, smem = 1024 + 2 * 65536;
  cudaFuncSetAttribute(k<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int issuers = 1; issuers <= 2; ++issuers) {
    k<N, MODE><<<148, 128, smem>>>(d, iters, issuers);
    cudaDeviceSynchronize();
    long long h[4];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const double n = iters * 4.0;
    printf("%s N=%3d issuers=%d: issue %.1f clk/MMA, retire %.1f clk/MMA per issuer (math %d clk)  %s\n", what, N, issuers,
           h[0] / n, h[1] / n, N / 2, cudaGetErrorString(cudaGetLastError()));
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  run<64, 0>(d, "SS");
  run<96, 0>(d, "SS");
  run<128, 0>(d, "SS");
  run<256, 0>(d, "SS");
  run<64, 2>(d, "SS-sw64");
  run<128, 2>(d, "SS-sw64");
  run<96, 3>(d, "TS-Bmn64");
  run<64, 1>(d, "TS");
  run<96, 1>(d, "TS");
  run<128, 1>(d, "TS");
  return 0;
}
