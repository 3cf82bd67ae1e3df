// sct_b200 — K4b row pass (vocab softmax-cross-entropy on a logits chunk) and K5 (discriminator head
// linears + GAN loss terms).
//
//   ce_rows       F.cross_entropy(logits, target, reduction='mean') without ignore_index
//                 (SCT-GAN/train.py:324) fused with its backward: one pass for (max, sum-exp, target
//                 logit), one pass that overwrites the chunk with d(loss)/d(logits).  The chunk is a
//                 bf16 scratch [rows, ld] written by the vocab GEMM; full [B*T, V] logits never exist.
//   small_linear  nn.Linear on [B, K] rows (B = batch) for disc_grammar_projection after pooling,
//                 disc_feature_extractor, disc_synthetic_head (model.py:250-271, 1190-1199): far too
//                 small for tensor cores, done in fp32 with one warp per weight row (no atomics).
//   gan_loss      BCE-with-logits vs ones, mean sigmoid confidence, the 0.3 / 0.8 branches and their
//                 penalties (train.py:1201-1234) with device-side predicates (no .item()).
#include "../../include/sct_b200.h"
#include "common.cuh"

namespace sct {
namespace {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ void online_merge(float& m, float& s, float m2, float s2) {
  const float mn = fmaxf(m, m2);
  if (mn == -INFINITY) {
    m = mn;
    s = 0.f;
    return;
  }
  s = s * exp2f(m - mn) + s2 * exp2f(m2 - mn);
  m = mn;
}

// one block (256 threads) per row
__global__ void __launch_bounds__(256)
ce_rows_kernel(__nv_bfloat16* __restrict__ logits, const int64_t* __restrict__ targets,
               float* __restrict__ row_loss, float* __restrict__ row_lse, int V, long long ld,
               float grad_scale, int write_grad) {
  __shared__ float sm_m[8], sm_s[8];
  __shared__ float sm_lse2;
  const int row = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __nv_bfloat16* x = logits + (long long)row * ld;
  const long long tgt = targets[row];
  const int V8 = V & ~7;
  if (tgt < 0) {  // excluded row (last position of each sequence): zero loss, zero gradient
    if (tid == 0) {
      row_loss[row] = 0.f;
      if (row_lse) row_lse[row] = 0.f;
    }
    if (write_grad) {
      for (int c = tid * 8; c < V8; c += 256 * 8) *reinterpret_cast<uint4*>(x + c) = make_uint4(0, 0, 0, 0);
      for (int c = V8 + tid; c < V; c += 256) x[c] = __float2bfloat16(0.f);
    }
    return;
  }
  if (tgt >= V) __trap();  // out-of-range class index: F.cross_entropy device-asserts here as well
  // pass 1: online (max, sum exp2) in the log2 domain
  float m = -INFINITY, s = 0.f;
  for (int c = tid * 8; c < V8; c += 256 * 8) {
    const uint4 u = *reinterpret_cast<const uint4*>(x + c);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    float v[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack_bf16(w[k]);
      v[2 * k] = f.x * kLog2e;
      v[2 * k + 1] = f.y * kLog2e;
    }
    float cm = v[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) cm = fmaxf(cm, v[k]);
    const float mn = fmaxf(m, cm);
    float cs = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) cs += exp2f(v[k] - mn);
    s = s * exp2f(m - mn) + cs;
    m = mn;
  }
  for (int c = V8 + tid; c < V; c += 256) {
    const float v = __bfloat162float(x[c]) * kLog2e;
    online_merge(m, s, v, 1.f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
    const float s2 = __shfl_xor_sync(0xffffffffu, s, o);
    online_merge(m, s, m2, s2);
  }
  if (lane == 0) {
    sm_m[warp] = m;
    sm_s[warp] = s;
  }
  __syncthreads();
  if (tid == 0) {
    float M = sm_m[0], S = sm_s[0];
    for (int w = 1; w < 8; ++w) online_merge(M, S, sm_m[w], sm_s[w]);
    const float lse2 = M + log2f(S);
    sm_lse2 = lse2;
    const float lt = __bfloat162float(x[tgt]);
    row_loss[row] = lse2 * kLn2 - lt;
    if (row_lse) row_lse[row] = lse2 * kLn2;
  }
  if (!write_grad) return;
  __syncthreads();
  const float lse2 = sm_lse2;
  // pass 2: d loss / d logit = (softmax - onehot) * grad_scale, overwriting the chunk
  for (int c = tid * 8; c < V8; c += 256 * 8) {
    const uint4 u = *reinterpret_cast<const uint4*>(x + c);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack_bf16(w[k]);
      float g0 = exp2f(f.x * kLog2e - lse2), g1 = exp2f(f.y * kLog2e - lse2);
      if (c + 2 * k == tgt) g0 -= 1.f;
      if (c + 2 * k + 1 == tgt) g1 -= 1.f;
      o[k] = pack_bf16(g0 * grad_scale, g1 * grad_scale);
    }
    *reinterpret_cast<uint4*>(x + c) = make_uint4(o[0], o[1], o[2], o[3]);
  }
  for (int c = V8 + tid; c < V; c += 256) {
    float g = exp2f(__bfloat162float(x[c]) * kLog2e - lse2);
    if (c == tgt) g -= 1.f;
    x[c] = __float2bfloat16(g * grad_scale);
  }
}

// ---------------------------------------------------------------------------------------------
// Row-resident variant (the training path: gradient requested and the row fits in shared memory).  The row is pulled
// into shared memory ONCE by a 1-D bulk copy (100 KB at V = 50265; two rows resident per SM, so one block's copy
// overlaps the other block's arithmetic) and all three sweeps read it from there: HBM sees one read and one write of
// the chunk instead of two reads and one write.  One exp2 per logit: sweep 1 = row maximum, sweep 2 = e_i =
// exp2(x_i - max) summed in fp32 and parked back in shared memory as bf16, sweep 3 = (e_i / sum - onehot) * grad_scale
// written to global memory.
// ---------------------------------------------------------------------------------------------
constexpr int kCeThreads = 512;
__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();  // `red` may still be read from the previous reduction
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < kCeThreads / 32; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

__global__ void __launch_bounds__(kCeThreads, 2)
ce_rows_resident_kernel(__nv_bfloat16* __restrict__ logits, const int64_t* __restrict__ targets,
                        float* __restrict__ row_loss, float* __restrict__ row_lse, int V, long long ld,
                        float grad_scale) {
  extern __shared__ __align__(16) uint8_t ce_smem[];
  __shared__ float red[kCeThreads / 32];
  __shared__ __align__(8) unsigned long long bar_storage;
  const int row = blockIdx.x;
  const int tid = threadIdx.x;
  __nv_bfloat16* x = logits + (long long)row * ld;
  const long long tgt = targets[row];
  const int V8 = V & ~7, nvec = V8 >> 3;
  if (tgt < 0) {  // excluded row (last position of each sequence): zero loss, zero gradient
    if (tid == 0) {
      row_loss[row] = 0.f;
      if (row_lse) row_lse[row] = 0.f;
    }
    for (int c = tid; c < nvec; c += kCeThreads) reinterpret_cast<uint4*>(x)[c] = make_uint4(0, 0, 0, 0);
    for (int c = V8 + tid; c < V; c += kCeThreads) x[c] = __float2bfloat16(0.f);
    return;
  }
  if (tgt >= V) __trap();  // out-of-range class index: F.cross_entropy device-asserts here as well
  const uint32_t bar = smem_u32(&bar_storage);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
    if (nvec > 0) {
      mbar_expect_tx(bar, (uint32_t)V8 * 2u);
      bulk_load_1d(smem_u32(ce_smem), x, (uint32_t)V8 * 2u, bar);
    } else {
      mbar_arrive(bar);
    }
  }
  // the (V % 8) tail elements and the target logit come straight from global memory, before anything is overwritten
  const int tcol = V8 + tid;
  const float tail = tcol < V ? __bfloat162float(x[tcol]) : -INFINITY;
  const float x_tgt = __bfloat162float(x[tgt]);
  __syncthreads();  // barrier initialised before anybody waits on it
  mbar_wait(bar, 0);
  uint4* sv = reinterpret_cast<uint4*>(ce_smem);
  // sweep 1: raw maximum (bf16 -> fp32 is a shift; log2e > 0 keeps the order)
  float m = tail;
  for (int c = tid; c < nvec; c += kCeThreads) {
    const uint4 u = sv[c];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) m = fmaxf(m, fmaxf(__uint_as_float(w[k] << 16), __uint_as_float(w[k] & 0xffff0000u)));
  }
  const float m2 = block_reduce(m, red, true) * kLog2e;  // row maximum in the log2 domain
  // sweep 2: e_i = exp2(x_i log2e - m2); fp32 sum; bf16(e_i) back into the row buffer
  const float2 sc2 = make_float2(kLog2e, kLog2e), nm2 = make_float2(-m2, -m2);
  float2 s2 = make_float2(0.f, 0.f);
  for (int c = tid; c < nvec; c += kCeThreads) {
    const uint4 u = sv[c];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 a = __ffma2_rn(make_float2(__uint_as_float(w[k] << 16), __uint_as_float(w[k] & 0xffff0000u)), sc2, nm2);
      const float2 e = make_float2(exp2f(a.x), exp2f(a.y));
      s2 = __fadd2_rn(s2, e);
      o[k] = pack_bf16(e.x, e.y);
    }
    sv[c] = make_uint4(o[0], o[1], o[2], o[3]);
  }
  const float e_tail = tcol < V ? exp2f(fmaf(tail, kLog2e, -m2)) : 0.f;
  const float S = block_reduce(s2.x + s2.y + e_tail, red, false);
  if (tid == 0) {
    const float lse2 = m2 + log2f(S);
    row_loss[row] = lse2 * kLn2 - x_tgt;
    if (row_lse) row_lse[row] = lse2 * kLn2;
  }
  // sweep 3: d loss / d logit = (softmax - onehot) * grad_scale, written over the chunk in global memory
  const float k = grad_scale / S;
  const float2 k2 = make_float2(k, k);
  const int tvec = (int)(tgt >> 3), tsub = (int)(tgt & 7);
  uint4* gx = reinterpret_cast<uint4*>(x);
  for (int c = tid; c < nvec; c += kCeThreads) {  // (each thread re-reads exactly the vectors it wrote in sweep 2)
    const uint4 u = sv[c];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    float g[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 v = __fmul2_rn(make_float2(__uint_as_float(w[q] << 16), __uint_as_float(w[q] & 0xffff0000u)), k2);
      g[2 * q] = v.x;
      g[2 * q + 1] = v.y;
    }
    if (c == tvec) {
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (q == tsub) g[q] -= grad_scale;
    }
    gx[c] = make_uint4(pack_bf16(g[0], g[1]), pack_bf16(g[2], g[3]), pack_bf16(g[4], g[5]), pack_bf16(g[6], g[7]));
  }
  if (tcol < V) {
    float gt = e_tail * k;
    if (tcol == tgt) gt -= grad_scale;
    x[tcol] = __float2bfloat16(gt);
  }
}

// ---------------------------------------------------------------------------------------------
// small linear (fp32 math): y[m, n] = sum_k x[m, k] * w[n, k] + b[n],  K = NV * 128
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld_row4(const float* xf, const __nv_bfloat16* xb, long long off) {
  if (xf) return *reinterpret_cast<const float4*>(xf + off);
  const uint2 u = *reinterpret_cast<const uint2*>(xb + off);
  const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}

template <int NV>
__global__ void __launch_bounds__(256)
small_linear_fwd_kernel(const float* __restrict__ xf, const __nv_bfloat16* __restrict__ xb,
                        const float* __restrict__ w, const float* __restrict__ bias,
                        float* __restrict__ yf, __nv_bfloat16* __restrict__ yb, int M, int N) {
  constexpr int K = NV * 128;
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (n >= N) return;
  float4 wr[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) wr[j] = *reinterpret_cast<const float4*>(w + (long long)n * K + (j * 32 + lane) * 4);
  const float bn = bias ? bias[n] : 0.f;
  // four rows per trip: their loads and their butterfly reductions are independent, so the ~600-cycle round trips of a
  // row's x loads and the five dependent shuffles of its reduction overlap (the one-row loop was a serial chain of both)
  for (int m0 = 0; m0 < M; m0 += 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int m = min(m0 + q, M - 1);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float4 xv = ld_row4(xf, xb, (long long)m * K + (j * 32 + lane) * 4);
        acc[q] += (xv.x * wr[j].x + xv.y * wr[j].y) + (xv.z * wr[j].z + xv.w * wr[j].w);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
    }
    if (lane < 4 && m0 + lane < M) {
      const float y = (lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3]) + bn;
      const int m = m0 + lane;
      if (yf) yf[(long long)m * N + n] = y;
      if (yb) yb[(long long)m * N + n] = __float2bfloat16(y);
    }
  }
}

// dw[n, :] = sum_m dy[m, n] * x[m, :];  db[n] = sum_m dy[m, n]
template <int NV>
__global__ void __launch_bounds__(256)
small_linear_bwd_w_kernel(const float* __restrict__ dyf, const __nv_bfloat16* __restrict__ dyb,
                          const float* __restrict__ xf, const __nv_bfloat16* __restrict__ xb,
                          float* __restrict__ dw, float* __restrict__ db, int M, int N) {
  constexpr int K = NV * 128;
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (n >= N) return;
  float4 acc[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  float sb = 0.f;
#pragma unroll 4
  for (int m = 0; m < M; ++m) {
    const float g = dyf ? dyf[(long long)m * N + n] : __bfloat162float(dyb[(long long)m * N + n]);
    sb += g;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float4 xv = ld_row4(xf, xb, (long long)m * K + (j * 32 + lane) * 4);
      acc[j].x += g * xv.x; acc[j].y += g * xv.y; acc[j].z += g * xv.z; acc[j].w += g * xv.w;
    }
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) *reinterpret_cast<float4*>(dw + (long long)n * K + (j * 32 + lane) * 4) = acc[j];
  if (lane == 0 && db) db[n] = sb;
}

// dx[m, 128j .. 128j+127] = sum_n dy[m, n] * w[n, 128j ..]: one block per (m, 128-column chunk); the 8 warps
// split the n loop (independent 512-byte row reads, 4 in flight per warp) and meet in shared memory.
template <int NV>
__global__ void __launch_bounds__(256)
small_linear_bwd_x_kernel(const float* __restrict__ dyf, const __nv_bfloat16* __restrict__ dyb,
                          const float* __restrict__ w, float* __restrict__ dxf,
                          __nv_bfloat16* __restrict__ dxb, int M, int N) {
  constexpr int K = NV * 128;
  __shared__ float4 red[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m = blockIdx.x, j = blockIdx.y;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int n = warp; n < N; n += 8) {
    const float g = dyf ? dyf[(long long)m * N + n] : __bfloat162float(dyb[(long long)m * N + n]);
    const float4 wv = *reinterpret_cast<const float4*>(w + (long long)n * K + (j * 32 + lane) * 4);
    acc.x += g * wv.x; acc.y += g * wv.y; acc.z += g * wv.z; acc.w += g * wv.w;
  }
  red[warp][lane] = acc;
  __syncthreads();
  if (warp == 0) {
    float4 s = red[0][lane];
#pragma unroll
    for (int q = 1; q < 8; ++q) {
      const float4 t = red[q][lane];
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    const long long o = (long long)m * K + (j * 32 + lane) * 4;
    if (dxf) *reinterpret_cast<float4*>(dxf + o) = s;
    if (dxb) {
      uint2 u;
      u.x = pack_bf16(s.x, s.y);
      u.y = pack_bf16(s.z, s.w);
      *reinterpret_cast<uint2*>(dxb + o) = u;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// GAN loss terms (train.py:1201-1234), single block.
//   out[0] = d_loss = mean softplus(-z) [+ mean s^2 + 2 mean s^4 if c > 0.8]
//   out[1] = adv    = mean softplus(z) if c < 0.3 else 0
//   out[2] = c      = mean sigmoid(z)   (or *c_in when given: the all-reduced global confidence)
//   out[3] = local sum of sigmoid(z)    (for the data-parallel all-reduce)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float softplusf(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__device__ float block_sum_256(float v, float* sm) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < 8; ++w) t += sm[w];
  return t;
}

__global__ void __launch_bounds__(256)
gan_loss_fwd_kernel(const float* __restrict__ z, int B, const float* __restrict__ c_in,
                    float* __restrict__ out) {
  __shared__ float sm[8];
  float s_sp_neg = 0.f, s_sp_pos = 0.f, s_sig = 0.f, s_sig2 = 0.f, s_sig4 = 0.f;
  for (int i = threadIdx.x; i < B; i += 256) {
    const float x = z[i];
    const float sg = sigmoidf_(x);
    s_sp_neg += softplusf(-x);
    s_sp_pos += softplusf(x);
    s_sig += sg;
    s_sig2 += sg * sg;
    s_sig4 += sg * sg * sg * sg;
  }
  s_sp_neg = block_sum_256(s_sp_neg, sm);
  s_sp_pos = block_sum_256(s_sp_pos, sm);
  s_sig = block_sum_256(s_sig, sm);
  s_sig2 = block_sum_256(s_sig2, sm);
  s_sig4 = block_sum_256(s_sig4, sm);
  if (threadIdx.x == 0) {
    const float inv = 1.f / B;
    const float c = c_in ? *c_in : s_sig * inv;
    float d = s_sp_neg * inv;
    if (c > 0.8f) d += 1.0f * s_sig2 * inv + 2.0f * s_sig4 * inv;
    out[0] = d;
    out[1] = (c < 0.3f) ? s_sp_pos * inv : 0.f;
    out[2] = c;
    out[3] = s_sig;
  }
}

__global__ void __launch_bounds__(256)
gan_loss_bwd_kernel(const float* __restrict__ z, int B, const float* __restrict__ c_ptr,
                    const float* __restrict__ g_d, const float* __restrict__ g_adv,
                    float* __restrict__ dz) {
  const float c = *c_ptr;
  const float gd = g_d ? *g_d : 0.f, ga = g_adv ? *g_adv : 0.f;
  const float inv = 1.f / B;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < B; i += gridDim.x * 256) {
    const float sg = sigmoidf_(z[i]);
    const float dsg = sg * (1.f - sg);
    float g = gd * (sg - 1.f) * inv;
    if (c > 0.8f) g += gd * (2.f * sg * dsg + 2.f * 4.f * sg * sg * sg * dsg) * inv;
    if (c < 0.3f) g += ga * sg * inv;
    dz[i] = g;
  }
}

}  // namespace
}  // namespace sct

using namespace sct;

#define DISPATCH_K(k, ...)                                               \
  switch (k) {                                                           \
    case 384: { constexpr int NV = 3; __VA_ARGS__; } break;              \
    case 768: { constexpr int NV = 6; __VA_ARGS__; } break;              \
    case 1536: { constexpr int NV = 12; __VA_ARGS__; } break;            \
    default: SCT_CHECK(false, "unsupported K=%d for small_linear (supported: 384, 768, 1536)", (int)(k)); \
  }

extern "C" {

int32_t sct_ce_rows(void* logits, const int64_t* targets, float* row_loss, float* row_lse,
                    int64_t rows, int64_t V, int64_t ld, float grad_scale, int32_t write_grad,
                    void* stream) {
  SCT_CHECK(logits && targets && row_loss, "null pointer");
  SCT_CHECK(ld % 8 == 0 && ld >= V, "logits pitch must be a multiple of 8 and >= V");
  SCT_CHECK(rows > 0 && V > 0, "empty input");
  const size_t row_bytes = (size_t)(V & ~7) * 2 + 16;
  if (write_grad && row_bytes <= 110 * 1024 && env_int("SCT_CE_RESIDENT", 1)) {  // two rows resident per SM
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(ce_rows_resident_kernel), (int)row_bytes)) return rc;
    ce_rows_resident_kernel<<<(unsigned)rows, kCeThreads, row_bytes, (cudaStream_t)stream>>>(
        (__nv_bfloat16*)logits, targets, row_loss, row_lse, (int)V, ld, grad_scale);
  } else {
    ce_rows_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(
        (__nv_bfloat16*)logits, targets, row_loss, row_lse, (int)V, ld, grad_scale, write_grad);
  }
  SCT_LAUNCH_CHECK();
  return 0;
}

int32_t sct_small_linear_fwd(const float* x_f32, const void* x_bf16, const float* w, const float* bias,
                             float* y_f32, void* y_bf16, int64_t M, int64_t N, int64_t K,
                             void* stream) {
  SCT_CHECK((x_f32 != nullptr) != (x_bf16 != nullptr), "exactly one input");
  SCT_CHECK(w && (y_f32 || y_bf16), "null pointer");
  const int blocks = (int)((N + 7) / 8);
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_K(K, (small_linear_fwd_kernel<NV><<<blocks, 256, 0, st>>>(
                    x_f32, (const __nv_bfloat16*)x_bf16, w, bias, y_f32, (__nv_bfloat16*)y_bf16, (int)M, (int)N)));
  SCT_LAUNCH_CHECK();
  return 0;
}

int32_t sct_small_linear_bwd(const float* dy_f32, const void* dy_bf16, const float* x_f32,
                             const void* x_bf16, const float* w, float* dx_f32, void* dx_bf16,
                             float* dw, float* db, int64_t M, int64_t N, int64_t K, void* stream) {
  SCT_CHECK((dy_f32 != nullptr) != (dy_bf16 != nullptr), "exactly one upstream gradient");
  SCT_CHECK((x_f32 != nullptr) != (x_bf16 != nullptr), "exactly one input");
  SCT_CHECK(w && dw, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  {
    const int blocks = (int)((N + 7) / 8);
    DISPATCH_K(K, (small_linear_bwd_w_kernel<NV><<<blocks, 256, 0, st>>>(
                      dy_f32, (const __nv_bfloat16*)dy_bf16, x_f32, (const __nv_bfloat16*)x_bf16, dw, db,
                      (int)M, (int)N)));
    SCT_LAUNCH_CHECK();
  }
  if (dx_f32 || dx_bf16) {
    DISPATCH_K(K, (small_linear_bwd_x_kernel<NV><<<dim3((unsigned)M, NV), 256, 0, st>>>(
                      dy_f32, (const __nv_bfloat16*)dy_bf16, w, dx_f32, (__nv_bfloat16*)dx_bf16, (int)M, (int)N)));
    SCT_LAUNCH_CHECK();
  }
  return 0;
}

int32_t sct_gan_loss_fwd(const float* z, int64_t B, const float* c_in, float* out4, void* stream) {
  SCT_CHECK(z && out4 && B > 0, "null pointer / empty batch");
  gan_loss_fwd_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(z, (int)B, c_in, out4);
  SCT_LAUNCH_CHECK();
  return 0;
}

int32_t sct_gan_loss_bwd(const float* z, int64_t B, const float* c, const float* g_d,
                         const float* g_adv, float* dz, void* stream) {
  SCT_CHECK(z && c && dz && B > 0, "null pointer / empty batch");
  gan_loss_bwd_kernel<<<(unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(z, (int)B, c, g_d,
                                                                                     g_adv, dz);
  SCT_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
