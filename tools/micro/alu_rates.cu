// Issue-rate microbenchmark for the softmax instruction mix of the attention kernels (sm_100a):
// FFMA vs FFMA2 (packed fp32x2), FADD2, FMNMX vs FMNMX3, MUFU.EX2, LOP3, IMAD, PRMT, F2FP pack, SEL.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o alu_rates alu_rates.cu && ./alu_rates
// Prints thread-level results per clock per SM (148 blocks x 512 threads, 8 independent chains per thread).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>

constexpr int ITERS = 2048;
constexpr int CH = 8;

template <int OP>
__global__ void __launch_bounds__(512) k(float* out, float seed, unsigned long long* clk) {
  float a[CH], b[CH];
  unsigned u[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    a[i] = seed + i + threadIdx.x;
    b[i] = seed * 0.5f + i;
    u[i] = (unsigned)(threadIdx.x * 977 + i) | 1u;
  }
  const float c0 = seed * 1.0001f, c1 = seed * 0.999f;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; i += 2) {
      if (OP == 0) {  // FFMA x2
        a[i] = fmaf(a[i], c0, c1);
        a[i + 1] = fmaf(a[i + 1], c0, c1);
      } else if (OP == 1) {  // FFMA2
        float2 r = __ffma2_rn(make_float2(a[i], a[i + 1]), make_float2(c0, c0), make_float2(c1, c1));
        a[i] = r.x;
        a[i + 1] = r.y;
      } else if (OP == 2) {  // MUFU.EX2 x2
        a[i] = exp2f(a[i]);
        a[i + 1] = exp2f(a[i + 1]);
      } else if (OP == 3) {  // FMNMX x2 (2 inputs)
        a[i] = fmaxf(a[i], b[i]);
        a[i + 1] = fmaxf(a[i + 1], b[i + 1]);
        b[i] += 1.f;  // keeps the compiler from hoisting (counted: +1 FADD per 2 max)
      } else if (OP == 4) {  // FMNMX3 (one instruction covers 2 new values)
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(b[i + 1]));
        b[i] += 1.f;
      } else if (OP == 5) {  // FADD2
        float2 r = __fadd2_rn(make_float2(a[i], a[i + 1]), make_float2(c0, c1));
        a[i] = r.x;
        a[i + 1] = r.y;
      } else if (OP == 6) {  // LOP3 x2
        u[i] = (u[i] & u[i + 1]) | (~u[i] & 0x5bd1e995u);
        u[i + 1] = (u[i + 1] ^ u[i]) & 0xdeadbeefu | u[i + 1];
      } else if (OP == 7) {  // IMAD x2
        u[i] = u[i] * 0x2C1B3C6Du + 12345u;
        u[i + 1] = u[i + 1] * 0x2C1B3C6Du + 54321u;
      } else if (OP == 8) {  // cvt pack bf16x2 (one per 2 values)
        __nv_bfloat162 v = __floats2bfloat162_rn(a[i], a[i + 1]);
        u[i] ^= *reinterpret_cast<unsigned*>(&v);
        a[i] += 1.f;
      } else if (OP == 9) {  // PRMT x2
        u[i] = __byte_perm(u[i], u[i + 1], 0x9988);
        u[i + 1] = __byte_perm(u[i + 1], u[i], 0xbbaa);
      } else if (OP == 10) {  // MUFU + FFMA2 + FADD2 + pack (the forward softmax core per 2 scores, no dropout)
        float2 r = __ffma2_rn(make_float2(a[i], a[i + 1]), make_float2(c0, c0), make_float2(c1, c1));
        const float e0 = exp2f(r.x), e1 = exp2f(r.y);
        float2 s = __fadd2_rn(make_float2(b[i], b[i + 1]), make_float2(e0, e1));
        b[i] = s.x;
        b[i + 1] = s.y;
        __nv_bfloat162 v = __floats2bfloat162_rn(e0, e1);
        u[i] ^= *reinterpret_cast<unsigned*>(&v);
      }
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
  unsigned ua = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    acc += a[i] + b[i];
    ua ^= u[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (float)ua;
  if (threadIdx.x == 0) clk[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <int OP>
void run(const char* name, double per_iter_results) {
  float* out;
  unsigned long long* clk;
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&clk, 148 * 8);
  k<OP><<<148, 512>>>(out, 1.0001f, clk);
  k<OP><<<148, 512>>>(out, 1.0001f, clk);
  cudaDeviceSynchronize();
  unsigned long long h[148];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += (double)h[i];
  avg /= 148;
  const double results = (double)ITERS * per_iter_results * 512;  // thread-level results per SM
  printf("%-44s %8.1f results/clk/SM   (%.0f clk)\n", name, results / avg, avg);
  cudaFree(out);
  cudaFree(clk);
}

int main() {
  run<0>("FFMA (scalar)", CH);
  run<1>("FFMA2 (two fp32 per instruction)", CH);
  run<2>("MUFU.EX2", CH);
  run<3>("FMNMX 2-input (+0.5 FADD)", CH);
  run<4>("FMNMX3 (2 new values / instr, +1 FADD)", CH);
  run<5>("FADD2", CH);
  run<6>("LOP3", CH);
  run<7>("IMAD", CH);
  run<8>("F2FP pack bf16x2 (values packed, +0.5 FADD)", CH);
  run<9>("PRMT", CH);
  run<10>("softmax core: FFMA2+2 MUFU+FADD2+pack per 2", CH);
  return 0;
}
