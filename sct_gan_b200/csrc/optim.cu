// sct_b200 — fused optimiser tail of the train step: the three clip_grad_norm_ calls, the
// per-parameter .item() norm loop, the NaN / norm > 1000 skip rule and AdamW with the four learning-rate
// groups of SCT-GAN/train.py:512-540, 1277-1311, as three multi-tensor launches without any host decision:
//   1. opt_sqnorm       sum of squares of every gradient chunk, then ONE block adds the chunk sums per clip scope
//                       (rest / disc_ / vuln heads) in a fixed order: no float atomics, so the clip coefficients — and
//                       with them every updated weight — are bit-identical on every data-parallel replica and from
//                       run to run (an atomicAdd version let replicas drift apart by ulps per step)
//   2. opt_clip_adamw   clip coefficients from those three numbers (the scopes nest: the second and third clip see
//                       gradients already scaled by the first), skip predicate, AdamW on fp32 master weights
//   3. opt_finish       per-tensor step counters, reported gradient norm and the stepped flag
// HBM-bound: 4 B/param for pass 1, 28 B/param (read g, p, m, v; write p, m, v) for pass 2, + 2 B/param for the bf16
// weight shadow the next forward's GEMMs read (written here instead of re-cast from the fp32 weights: -4 B/param).
#include <math.h>

#include "../../include/sct_b200.h"
#include "common.cuh"

namespace sct {
namespace {

constexpr int kChunk = 32768;  // elements per block
constexpr int kThreads = 256;

struct OptTensor {  // mirrors sct_opt_tensor in the header (80 bytes)
  float* p;
  const float* g;
  float* m;
  float* v;
  float* step;
  long long numel;
  float lr;
  float wd;
  int seg;
  int pad;
  __nv_bfloat16* shadow;  // nullable: bf16 copy of p for the tensor-core GEMMs, refreshed with the update
  long long pad2;
};
static_assert(sizeof(OptTensor) == 80, "table layout");

struct ClipInfo {
  float c_rest, c_disc, c_vuln, total;
  bool skip;
};

__device__ __forceinline__ ClipInfo clip_info(const float* sq, const float* loss, float max_norm, float disc_mult,
                                              float vuln_mult) {
  ClipInfo ci;
  const float n_all = sqrtf(sq[0] + sq[1] + sq[2]);
  const float c1 = fminf(1.0f, max_norm / (n_all + 1e-6f));                       // clip_grad_norm_(all, max)
  const float c2 = fminf(1.0f, max_norm * disc_mult / (c1 * sqrtf(sq[1]) + 1e-6f));  // then the disc_ scope
  const float c3 = fminf(1.0f, max_norm * vuln_mult / (c1 * sqrtf(sq[2]) + 1e-6f));  // then the vulnerability heads
  ci.c_rest = c1;
  ci.c_disc = c1 * c2;
  ci.c_vuln = c1 * c3;
  ci.total = sqrtf(c1 * c1 * (sq[0] + c2 * c2 * sq[1] + c3 * c3 * sq[2]));
  const float l = loss ? *loss : 0.f;
  ci.skip = !(isfinite(l)) || !(isfinite(ci.total)) || ci.total > 1000.0f;
  return ci;
}

// 16-byte vector path only when the element count AND every pointer allow it (views into larger storages, e.g. a
// parameter that is a slice of a flat buffer, may be only 4-byte aligned)
__device__ __forceinline__ bool vec4_ok(const OptTensor& t) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) |
                      reinterpret_cast<uintptr_t>(t.m) | reinterpret_cast<uintptr_t>(t.v);
  return (t.numel & 3) == 0 && (a & 15) == 0 && (reinterpret_cast<uintptr_t>(t.shadow) & 7) == 0;
}

__global__ void __launch_bounds__(kThreads)
opt_sqnorm_kernel(const OptTensor* __restrict__ tab, const int2* __restrict__ chunks, float* __restrict__ part) {
  __shared__ float red[kThreads / 32];
  const int2 ck = chunks[blockIdx.x];
  const OptTensor t = tab[ck.x];
  const long long base = (long long)ck.y * kChunk;
  const long long end = min(base + kChunk, t.numel);
  float s = 0.f;
  if (vec4_ok(t)) {
    for (long long i = base + threadIdx.x * 4; i < end; i += kThreads * 4) {
      const float4 g = *reinterpret_cast<const float4*>(t.g + i);
      s += (g.x * g.x + g.y * g.y) + (g.z * g.z + g.w * g.w);
    }
  } else {
    for (long long i = base + threadIdx.x; i < end; i += kThreads) s += t.g[i] * t.g[i];
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) tot += red[w];
    part[blockIdx.x] = tot;
  }
}

constexpr int kReduceThreads = 1024;
__global__ void __launch_bounds__(kReduceThreads)
opt_sqnorm_reduce_kernel(const OptTensor* __restrict__ tab, const int2* __restrict__ chunks,
                         const float* __restrict__ part, int n_chunks, float* __restrict__ sq) {
  __shared__ double red[3][kReduceThreads];
  double s[3] = {0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < n_chunks; i += kReduceThreads) s[tab[chunks[i].x].seg] += (double)part[i];
#pragma unroll
  for (int k = 0; k < 3; ++k) red[k][threadIdx.x] = s[k];
  __syncthreads();
  for (int off = kReduceThreads / 2; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) {
#pragma unroll
      for (int k = 0; k < 3; ++k) red[k][threadIdx.x] += red[k][threadIdx.x + off];
    }
    __syncthreads();
  }
  if (threadIdx.x < 3) sq[threadIdx.x] = (float)red[threadIdx.x][0];
}

__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, float lr, float wd, float b1, float b2,
                                          float eps, float inv_bias1, float inv_sqrt_bias2) {
  p *= 1.0f - lr * wd;
  m = b1 * m + (1.0f - b1) * g;
  v = b2 * v + (1.0f - b2) * g * g;
  const float denom = sqrtf(v) * inv_sqrt_bias2 + eps;
  p -= lr * inv_bias1 * m / denom;
}

__global__ void __launch_bounds__(kThreads)
opt_clip_adamw_kernel(const OptTensor* __restrict__ tab, const int2* __restrict__ chunks, const float* __restrict__ sq,
                      const float* __restrict__ loss, float max_norm, float disc_mult, float vuln_mult, float b1,
                      float b2, float eps) {
  const ClipInfo ci = clip_info(sq, loss, max_norm, disc_mult, vuln_mult);
  if (ci.skip) return;
  const int2 ck = chunks[blockIdx.x];
  const OptTensor t = tab[ck.x];
  const float coef = t.seg == 0 ? ci.c_rest : (t.seg == 1 ? ci.c_disc : ci.c_vuln);
  const float step = *t.step + 1.0f;
  const float inv_bias1 = 1.0f / (1.0f - powf(b1, step));
  const float inv_sqrt_bias2 = rsqrtf(1.0f - powf(b2, step));
  const long long base = (long long)ck.y * kChunk;
  const long long end = min(base + kChunk, t.numel);
  if (vec4_ok(t)) {
    for (long long i = base + threadIdx.x * 4; i < end; i += kThreads * 4) {
      float4 g = *reinterpret_cast<const float4*>(t.g + i);
      float4 p = *reinterpret_cast<float4*>(t.p + i);
      float4 m = *reinterpret_cast<float4*>(t.m + i);
      float4 v = *reinterpret_cast<float4*>(t.v + i);
      adamw_one(p.x, g.x * coef, m.x, v.x, t.lr, t.wd, b1, b2, eps, inv_bias1, inv_sqrt_bias2);
      adamw_one(p.y, g.y * coef, m.y, v.y, t.lr, t.wd, b1, b2, eps, inv_bias1, inv_sqrt_bias2);
      adamw_one(p.z, g.z * coef, m.z, v.z, t.lr, t.wd, b1, b2, eps, inv_bias1, inv_sqrt_bias2);
      adamw_one(p.w, g.w * coef, m.w, v.w, t.lr, t.wd, b1, b2, eps, inv_bias1, inv_sqrt_bias2);
      *reinterpret_cast<float4*>(t.p + i) = p;
      *reinterpret_cast<float4*>(t.m + i) = m;
      *reinterpret_cast<float4*>(t.v + i) = v;
      if (t.shadow != nullptr) {
        uint2 u;
        u.x = pack_bf16(p.x, p.y);
        u.y = pack_bf16(p.z, p.w);
        *reinterpret_cast<uint2*>(t.shadow + i) = u;
      }
    }
  } else {
    for (long long i = base + threadIdx.x; i < end; i += kThreads) {
      float p = t.p[i], m = t.m[i], v = t.v[i];
      adamw_one(p, t.g[i] * coef, m, v, t.lr, t.wd, b1, b2, eps, inv_bias1, inv_sqrt_bias2);
      t.p[i] = p;
      t.m[i] = m;
      t.v[i] = v;
      if (t.shadow != nullptr) t.shadow[i] = __float2bfloat16(p);
    }
  }
}

__global__ void __launch_bounds__(kThreads)
opt_finish_kernel(const OptTensor* __restrict__ tab, int n_tensors, const float* __restrict__ sq,
                  const float* __restrict__ loss, float max_norm, float disc_mult, float vuln_mult,
                  float* __restrict__ out2) {
  const ClipInfo ci = clip_info(sq, loss, max_norm, disc_mult, vuln_mult);
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i < n_tensors && !ci.skip) *tab[i].step += 1.0f;
  if (i == 0) {
    out2[0] = ci.total;
    out2[1] = ci.skip ? 0.f : 1.f;
  }
}

}  // namespace
}  // namespace sct

using namespace sct;

extern "C" {

int32_t sct_opt_chunk_elems(void) { return kChunk; }

int32_t sct_clip_adamw_step(const void* table, int32_t n_tensors, const void* chunks, int32_t n_chunks,
                            const float* loss, float* sqnorm3, float* out2, float max_norm, float disc_mult,
                            float vuln_mult, float beta1, float beta2, float eps, void* stream) {
  SCT_CHECK(table && chunks && sqnorm3 && out2, "null pointer");
  SCT_CHECK(n_tensors > 0 && n_chunks > 0, "empty parameter table");
  cudaStream_t st = (cudaStream_t)stream;
  const OptTensor* tab = static_cast<const OptTensor*>(table);
  const int2* ck = static_cast<const int2*>(chunks);
  float* part = sqnorm3 + 4;  // [n_chunks] per-chunk sums of squares
  opt_sqnorm_kernel<<<n_chunks, kThreads, 0, st>>>(tab, ck, part);
  SCT_LAUNCH_CHECK();
  opt_sqnorm_reduce_kernel<<<1, kReduceThreads, 0, st>>>(tab, ck, part, n_chunks, sqnorm3);
  SCT_LAUNCH_CHECK();
  opt_clip_adamw_kernel<<<n_chunks, kThreads, 0, st>>>(tab, ck, sqnorm3, loss, max_norm, disc_mult, vuln_mult, beta1,
                                                       beta2, eps);
  SCT_LAUNCH_CHECK();
  opt_finish_kernel<<<(n_tensors + kThreads - 1) / kThreads, kThreads, 0, st>>>(tab, n_tensors, sqnorm3, loss, max_norm,
                                                                               disc_mult, vuln_mult, out2);
  SCT_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
