// sct_b200 — K1 / K4a and the other HBM-bound row kernels of the SCT-GAN hot path.
//
// All of these are bandwidth kernels: one warp owns one row of d = NV*128 features, every lane moves
// 16-byte vectors (fully coalesced 512 B per warp per access), statistics are reduced with shuffles and
// kept in fp32.  Dropout masks are never stored: they are regenerated from (seed, offset, index).
//
//   K1   embed_ln_pe      nn.Embedding * sqrt(d) -> Dropout -> LayerNorm -> + pe[s]
//                          (SCT-GAN/model.py:412-421, 944-947, 8-21)
//   K4a  add_dropout_ln    x' = x + alpha*Dropout(branch); y = LayerNorm(x')   (pre-norm residual sites,
//                          torch nn/modules/transformer.py:944-983, 1131-1205; model.py:439, 451, 957)
//        ln_act            Linear -> LayerNorm -> GELU -> Dropout blocks (model.py:225-235, 253-271)
//        gelu_dropout      FFN activation (activation='gelu', exact erf; model.py:62, 74)
//        colsum            bias gradients
#include <stdlib.h>

#include "../../include/sct_b200.h"
#include "common.cuh"

namespace sct {
namespace {

constexpr int kWarpsPerBlock = 8;
constexpr float kLnEps = 1e-5f;

struct DropCfg {
  uint64_t seed, offset;
  const unsigned long long* epoch;  // device-resident epoch added to offset at run time (nullable)
  uint32_t thresh16;  // drop if u16 < thresh16
  float inv_keep;     // 1/(1-p)
};

__device__ __forceinline__ DropCfg resolve_epoch(DropCfg dc) {
  if (dc.thresh16 != 0 && dc.epoch != nullptr) dc.offset += *dc.epoch * 0x100000ull;
  return dc;
}

__host__ inline DropCfg make_drop(float p, uint64_t seed, uint64_t offset, const uint64_t* epoch) {
  DropCfg c;
  c.seed = seed;
  c.offset = offset;
  c.epoch = reinterpret_cast<const unsigned long long*>(epoch);
  double t = (double)p * 65536.0 + 0.5;
  if (t < 0) t = 0;
  if (t > 65535.0) t = 65535.0;
  c.thresh16 = p > 0.f ? (uint32_t)t : 0u;
  c.inv_keep = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  return c;
}

// dropout multipliers for 4 consecutive elements starting at index idx (idx % 4 == 0): one strong hash of
// (seed, offset, idx / 4) gives the first two 16-bit uniforms, a xorshift-multiply step of it the other two
__device__ __forceinline__ void drop4(const DropCfg& dc, uint64_t idx, float (&m)[4]) {
  if (dc.thresh16 == 0) {
    m[0] = m[1] = m[2] = m[3] = 1.f;
    return;
  }
  const uint32_t h0 = rng_pair(dc.seed, dc.offset, idx >> 2);
  uint32_t h1 = (h0 ^ (h0 >> 15)) * 0x2C1B3C6Du;
  h1 ^= h1 >> 13;
  m[0] = ((h0 & 0xFFFFu) >= dc.thresh16) ? dc.inv_keep : 0.f;
  m[1] = ((h0 >> 16) >= dc.thresh16) ? dc.inv_keep : 0.f;
  m[2] = ((h1 >> 16) >= dc.thresh16) ? dc.inv_keep : 0.f;
  m[3] = ((h1 & 0xFFFFu) >= dc.thresh16) ? dc.inv_keep : 0.f;
}
// 8 consecutive elements (idx % 8 == 0): one strong hash + three cheap steps
__device__ __forceinline__ void drop8(const DropCfg& dc, uint64_t idx, float (&m)[8]) {
  if (dc.thresh16 == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = 1.f;
    return;
  }
  uint32_t h = rng_pair(dc.seed, dc.offset ^ 0x9E3779B97F4A7C15ull, idx >> 3);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[2 * i] = ((h & 0xFFFFu) >= dc.thresh16) ? dc.inv_keep : 0.f;
    m[2 * i + 1] = ((h >> 16) >= dc.thresh16) ? dc.inv_keep : 0.f;
    h = (h ^ (h >> 15)) * 0x2C1B3C6Du;
    h ^= h >> 13;
  }
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 ldbf4(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void stbf4(__nv_bfloat16* p, float4 v) {
  uint2 u;
  u.x = pack_bf16(v.x, v.y);
  u.y = pack_bf16(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = u;
}

// Row statistics over NV float4 per lane (two-pass in registers -> exact biased variance).
template <int NV>
__device__ __forceinline__ void row_stats(const float4 (&x)[NV], float& mean, float& rstd) {
  constexpr float inv_d = 1.0f / (NV * 128);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) s += (x[j].x + x[j].y) + (x[j].z + x[j].w);
  mean = warp_sum(s) * inv_d;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const float a = x[j].x - mean, b = x[j].y - mean, c = x[j].z - mean, e = x[j].w - mean;
    q += (a * a + b * b) + (c * c + e * e);
  }
  rstd = rsqrtf(warp_sum(q) * inv_d + kLnEps);
}

// Block-level reduction of per-lane column partials into global fp32 accumulators (atomicAdd).
template <int NV>
__device__ __forceinline__ void flush_col_partials(float4 (&acc)[NV], float* __restrict__ gout,
                                                   float* smem /* [kWarpsPerBlock][NV*128] */) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < NV; ++j) st4(smem + warp * (NV * 128) + (j * 32 + lane) * 4, acc[j]);
  __syncthreads();
  for (int c = threadIdx.x; c < NV * 128; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarpsPerBlock; ++w) s += smem[w * (NV * 128) + c];
    atomicAdd(gout + c, s);
  }
}

// =================================================================================================
// K1 forward
// =================================================================================================
template <int NV>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
embed_ln_pe_fwd_kernel(const int64_t* __restrict__ ids, const float* __restrict__ table,
                       const float* __restrict__ gamma, const float* __restrict__ beta,
                       const float* __restrict__ pe, float* __restrict__ out_f32,
                       __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ stats, int n_tok,
                       int seq_len, int vocab, float scale, DropCfg dc_in) {
  const DropCfg dc = resolve_epoch(dc_in);
  constexpr int D = NV * 128;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= n_tok) return;
  long long id = ids[row];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const float* w = table + id * (long long)D;
  const int s = row % seq_len;
  float4 x[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (j * 32 + lane) * 4;
    float4 v = ld4(w + c);
    float m[4];
    drop4(dc, (uint64_t)row * D + c, m);
    x[j] = make_float4(v.x * scale * m[0], v.y * scale * m[1], v.z * scale * m[2], v.w * scale * m[3]);
  }
  float mean, rstd;
  row_stats<NV>(x, mean, rstd);
  if (lane == 0) {
    stats[2 * row] = mean;
    stats[2 * row + 1] = rstd;
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (j * 32 + lane) * 4;
    const float4 g = ld4(gamma + c), b = ld4(beta + c), e = ld4(pe + (long long)s * D + c);
    float4 y;
    y.x = (x[j].x - mean) * rstd * g.x + b.x + e.x;
    y.y = (x[j].y - mean) * rstd * g.y + b.y + e.y;
    y.z = (x[j].z - mean) * rstd * g.z + b.z + e.z;
    y.w = (x[j].w - mean) * rstd * g.w + b.w + e.w;
    if (out_f32) st4(out_f32 + (long long)row * D + c, y);
    if (out_bf16) stbf4(out_bf16 + (long long)row * D + c, y);
  }
}

// =================================================================================================
// K1 backward: LayerNorm backward -> dropout/scale -> scatter-add into the fp32 table gradient
// =================================================================================================
template <int NV>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
embed_ln_pe_bwd_kernel(const float* __restrict__ g_f32, const __nv_bfloat16* __restrict__ g_bf16,
                       const int64_t* __restrict__ ids, const float* __restrict__ table,
                       const float* __restrict__ gamma, const float* __restrict__ stats,
                       float* __restrict__ dtable, float* __restrict__ dgamma,
                       float* __restrict__ dbeta, int n_tok, int vocab, float scale, DropCfg dc_in) {
  const DropCfg dc = resolve_epoch(dc_in);
  constexpr int D = NV * 128;
  constexpr float inv_d = 1.0f / D;
  __shared__ float red[kWarpsPerBlock * D];
  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int warps_total = gridDim.x * kWarpsPerBlock;
  float4 ag[NV], ab[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) ag[j] = ab[j] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int row = warp_global; row < n_tok; row += warps_total) {
    long long id = ids[row];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const float* w = table + id * (long long)D;
    const float mean = stats[2 * row], rstd = stats[2 * row + 1];
    float4 xh[NV], g[NV], msk[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = (j * 32 + lane) * 4;
      const float4 v = ld4(w + c);
      float m[4];
      drop4(dc, (uint64_t)row * D + c, m);
      msk[j] = make_float4(m[0], m[1], m[2], m[3]);
      xh[j] = make_float4((v.x * scale * m[0] - mean) * rstd, (v.y * scale * m[1] - mean) * rstd,
                          (v.z * scale * m[2] - mean) * rstd, (v.w * scale * m[3] - mean) * rstd);
      float4 gy = make_float4(0.f, 0.f, 0.f, 0.f);  // either or both incoming gradients (fp32 residual stream, bf16 copy)
      if (g_f32) gy = ld4(g_f32 + (long long)row * D + c);
      if (g_bf16) {
        const float4 t = ldbf4(g_bf16 + (long long)row * D + c);
        gy.x += t.x; gy.y += t.y; gy.z += t.z; gy.w += t.w;
      }
      const float4 gm = ld4(gamma + c);
      ag[j].x += gy.x * xh[j].x; ag[j].y += gy.y * xh[j].y; ag[j].z += gy.z * xh[j].z; ag[j].w += gy.w * xh[j].w;
      ab[j].x += gy.x; ab[j].y += gy.y; ab[j].z += gy.z; ab[j].w += gy.w;
      g[j] = make_float4(gy.x * gm.x, gy.y * gm.y, gy.z * gm.z, gy.w * gm.w);
      s1 += (g[j].x + g[j].y) + (g[j].z + g[j].w);
      s2 += (g[j].x * xh[j].x + g[j].y * xh[j].y) + (g[j].z * xh[j].z + g[j].w * xh[j].w);
    }
    s1 = warp_sum(s1) * inv_d;
    s2 = warp_sum(s2) * inv_d;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = (j * 32 + lane) * 4;
      float4 dx;
      dx.x = rstd * (g[j].x - s1 - xh[j].x * s2) * scale * msk[j].x;
      dx.y = rstd * (g[j].y - s1 - xh[j].y * s2) * scale * msk[j].y;
      dx.z = rstd * (g[j].z - s1 - xh[j].z * s2) * scale * msk[j].z;
      dx.w = rstd * (g[j].w - s1 - xh[j].w * s2) * scale * msk[j].w;
      atomicAdd(reinterpret_cast<float4*>(dtable + id * (long long)D + c), dx);
    }
  }
  flush_col_partials<NV>(ag, dgamma, red);
  flush_col_partials<NV>(ab, dbeta, red);
}

// =================================================================================================
// K4a forward:  x' = x + alpha * dropout(branch);  y_ln = LN(x') (bf16);  y_cast = bf16(x')
// =================================================================================================
template <int NV>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
add_dropout_ln_fwd_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ branch,
                          float alpha, const float* __restrict__ gamma,
                          const float* __restrict__ beta, float* __restrict__ x_out,
                          __nv_bfloat16* __restrict__ y_ln, __nv_bfloat16* __restrict__ y_cast,
                          float* __restrict__ stats, int n_rows, DropCfg dc_in) {
  const DropCfg dc = resolve_epoch(dc_in);
  constexpr int D = NV * 128;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const long long base = (long long)row * D;
  float4 v[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (j * 32 + lane) * 4;
    v[j] = x ? ld4(x + base + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (branch) {
      const float4 b = ldbf4(branch + base + c);
      float m[4];
      drop4(dc, (uint64_t)base + c, m);
      v[j].x += alpha * b.x * m[0];
      v[j].y += alpha * b.y * m[1];
      v[j].z += alpha * b.z * m[2];
      v[j].w += alpha * b.w * m[3];
    }
    if (x_out) st4(x_out + base + c, v[j]);
    if (y_cast) stbf4(y_cast + base + c, v[j]);
  }
  if (y_ln) {
    float mean, rstd;
    row_stats<NV>(v, mean, rstd);
    if (lane == 0) {
      stats[2 * row] = mean;
      stats[2 * row + 1] = rstd;
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = (j * 32 + lane) * 4;
      const float4 g = ld4(gamma + c), b = ld4(beta + c);
      float4 y;
      y.x = (v[j].x - mean) * rstd * g.x + b.x;
      y.y = (v[j].y - mean) * rstd * g.y + b.y;
      y.z = (v[j].z - mean) * rstd * g.z + b.z;
      y.w = (v[j].w - mean) * rstd * g.w + b.w;
      stbf4(y_ln + base + c, y);
    }
  }
}

// Same arithmetic, rows staged through shared memory by 1-D bulk copies (TMA engine): every warp owns a ring of
// kLnSlots row buffers and keeps the next rows in flight while it works on the current one, so the bytes in flight
// no longer depend on how many warps happen to be in their load phase (the plain kernel reached 0.65 of HBM bandwidth:
// 42 % occupancy, long-scoreboard stalls).  Persistent: 2 blocks per SM, warps walk the rows with a grid stride.
constexpr int kLnSlots = 3;
template <int NV>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
add_dropout_ln_fwd_staged_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ branch, float alpha,
                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                 float* __restrict__ x_out, __nv_bfloat16* __restrict__ y_ln,
                                 __nv_bfloat16* __restrict__ y_cast, float* __restrict__ stats, int n_rows,
                                 DropCfg dc_in) {
  const DropCfg dc = resolve_epoch(dc_in);
  constexpr int D = NV * 128;
  constexpr uint32_t XB = D * 4, BB = D * 2, SLOT = XB + BB;
  extern __shared__ __align__(128) uint8_t ln_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t ring = smem_u32(ln_smem) + warp * (kLnSlots * SLOT);
  const uint32_t bars = smem_u32(ln_smem) + kWarpsPerBlock * kLnSlots * SLOT + warp * (kLnSlots * 8);
  const uint8_t* ring_gen = ln_smem + warp * (kLnSlots * SLOT);
  const int warp_global = blockIdx.x * kWarpsPerBlock + warp;
  const int warps_total = gridDim.x * kWarpsPerBlock;
  const uint32_t bytes = (x ? XB : 0u) + (branch ? BB : 0u);
  if (lane == 0) {
    for (int s = 0; s < kLnSlots; ++s) mbar_init(bars + 8 * s, 1);
    fence_mbar_init();
  }
  __syncwarp();
  auto issue = [&](int row, int slot) {  // lane 0
    const long long base = (long long)row * D;
    mbar_expect_tx(bars + 8 * slot, bytes);
    if (x) bulk_load_1d(ring + slot * SLOT, x + base, XB, bars + 8 * slot);
    if (branch) bulk_load_1d(ring + slot * SLOT + XB, branch + base, BB, bars + 8 * slot);
  };
  if (lane == 0)
    for (int s = 0; s < kLnSlots; ++s) {
      const int row = warp_global + s * warps_total;
      if (row < n_rows) issue(row, s);
    }
  int k = 0;
  for (int row = warp_global; row < n_rows; row += warps_total, ++k) {
    const int slot = k % kLnSlots;
    mbar_wait(bars + 8 * slot, (uint32_t)(k / kLnSlots) & 1u);
    const long long base = (long long)row * D;
    const float* xs = reinterpret_cast<const float*>(ring_gen + slot * SLOT);
    const __nv_bfloat16* bs = reinterpret_cast<const __nv_bfloat16*>(ring_gen + slot * SLOT + XB);
    float4 v[NV], bv[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = (j * 32 + lane) * 4;
      v[j] = x ? ld4(xs + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      bv[j] = branch ? ldbf4(bs + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();  // every lane has its copy of the row: the slot can take the row kLnSlots ahead
    if (lane == 0) {
      const int nxt = row + kLnSlots * warps_total;
      if (nxt < n_rows) {
        fence_proxy_async_smem();
        issue(nxt, slot);
      }
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = (j * 32 + lane) * 4;
      if (branch) {
        float m[4];
        drop4(dc, (uint64_t)base + c, m);
        v[j].x += alpha * bv[j].x * m[0];
        v[j].y += alpha * bv[j].y * m[1];
        v[j].z += alpha * bv[j].z * m[2];
        v[j].w += alpha * bv[j].w * m[3];
      }
      if (x_out) st4(x_out + base + c, v[j]);
      if (y_cast) stbf4(y_cast + base + c, v[j]);
    }
    if (y_ln) {
      float mean, rstd;
      row_stats<NV>(v, mean, rstd);
      if (lane == 0) {
        stats[2 * row] = mean;
        stats[2 * row + 1] = rstd;
      }
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c = (j * 32 + lane) * 4;
        const float4 g = ld4(gamma + c), b = ld4(beta + c);
        float4 y;
        y.x = (v[j].x - mean) * rstd * g.x + b.x;
        y.y = (v[j].y - mean) * rstd * g.y + b.y;
        y.z = (v[j].z - mean) * rstd * g.z + b.z;
        y.w = (v[j].w - mean) * rstd * g.w + b.w;
        stbf4(y_ln + base + c, y);
      }
    }
  }
}

// K4a backward:  g_x = g_xout + g_ycast + LN_bwd(g_yln);  g_branch = alpha * mask * g_x
template <int NV>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
add_dropout_ln_bwd_kernel(const float* __restrict__ g_xout, const __nv_bfloat16* __restrict__ g_yln,
                          const __nv_bfloat16* __restrict__ g_ycast,
                          const float* __restrict__ xprime, const float* __restrict__ stats,
                          const float* __restrict__ gamma, float alpha, float* __restrict__ g_x,
                          __nv_bfloat16* __restrict__ g_branch, float* __restrict__ dgamma,
                          float* __restrict__ dbeta, int n_rows, DropCfg dc_in) {
  const DropCfg dc = resolve_epoch(dc_in);
  constexpr int D = NV * 128;
  constexpr float inv_d = 1.0f / D;
  __shared__ float red[kWarpsPerBlock * D];
  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int warps_total = gridDim.x * kWarpsPerBlock;
  float4 ag[NV], ab[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) ag[j] = ab[j] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int row = warp_global; row < n_rows; row += warps_total) {
    const long long base = (long long)row * D;
    float4 acc[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = (j * 32 + lane) * 4;
      acc[j] = g_xout ? ld4(g_xout + base + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (g_ycast) {
        const float4 t = ldbf4(g_ycast + base + c);
        acc[j].x += t.x; acc[j].y += t.y; acc[j].z += t.z; acc[j].w += t.w;
      }
    }
    if (g_yln) {
      const float mean = stats[2 * row], rstd = stats[2 * row + 1];
      float4 xh[NV], g[NV];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c = (j * 32 + lane) * 4;
        const float4 xv = ld4(xprime + base + c);
        xh[j] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd,
                            (xv.w - mean) * rstd);
        const float4 gy = ldbf4(g_yln + base + c);
        const float4 gm = ld4(gamma + c);
        ag[j].x += gy.x * xh[j].x; ag[j].y += gy.y * xh[j].y; ag[j].z += gy.z * xh[j].z; ag[j].w += gy.w * xh[j].w;
        ab[j].x += gy.x; ab[j].y += gy.y; ab[j].z += gy.z; ab[j].w += gy.w;
        g[j] = make_float4(gy.x * gm.x, gy.y * gm.y, gy.z * gm.z, gy.w * gm.w);
        s1 += (g[j].x + g[j].y) + (g[j].z + g[j].w);
        s2 += (g[j].x * xh[j].x + g[j].y * xh[j].y) + (g[j].z * xh[j].z + g[j].w * xh[j].w);
      }
      s1 = warp_sum(s1) * inv_d;
      s2 = warp_sum(s2) * inv_d;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        acc[j].x += rstd * (g[j].x - s1 - xh[j].x * s2);
        acc[j].y += rstd * (g[j].y - s1 - xh[j].y * s2);
        acc[j].z += rstd * (g[j].z - s1 - xh[j].z * s2);
        acc[j].w += rstd * (g[j].w - s1 - xh[j].w * s2);
      }
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = (j * 32 + lane) * 4;
      if (g_x) st4(g_x + base + c, acc[j]);
      if (g_branch) {
        float m[4];
        drop4(dc, (uint64_t)base + c, m);
        stbf4(g_branch + base + c,
              make_float4(alpha * m[0] * acc[j].x, alpha * m[1] * acc[j].y, alpha * m[2] * acc[j].z,
                          alpha * m[3] * acc[j].w));
      }
    }
  }
  if (g_yln && dgamma) {
    flush_col_partials<NV>(ag, dgamma, red);
    flush_col_partials<NV>(ab, dbeta, red);
  }
}

// The common case of the backward (both the residual gradient and the LayerNorm gradient arrive, d = 768), software-
// pipelined by one row: the kernel above runs one block of 8 warps per SM (156 registers), and a warp has loads in
// flight only while it is not in its reduction / store phase.  Registers are free up to 255 at that occupancy, so the
// NEXT row's three inputs are fetched into a second register set before the current row's reductions.
template <int NV>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 1)
add_dropout_ln_bwd_pf_kernel(const float* __restrict__ g_xout, const __nv_bfloat16* __restrict__ g_yln,
                             const __nv_bfloat16* __restrict__ g_ycast, const float* __restrict__ xprime,
                             const float* __restrict__ stats, const float* __restrict__ gamma, float alpha,
                             float* __restrict__ g_x, __nv_bfloat16* __restrict__ g_branch,
                             float* __restrict__ dgamma, float* __restrict__ dbeta, int n_rows, DropCfg dc_in) {
  const DropCfg dc = resolve_epoch(dc_in);
  constexpr int D = NV * 128;
  constexpr float inv_d = 1.0f / D;
  __shared__ float red[kWarpsPerBlock * D];
  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int warps_total = gridDim.x * kWarpsPerBlock;
  float4 ag[NV], ab[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) ag[j] = ab[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 n_gx[NV], n_xp[NV];
  uint2 n_gy[NV];
  float n_mean = 0.f, n_rstd = 0.f;
  auto fetch = [&](int row) {
    const long long base = (long long)row * D;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = (j * 32 + lane) * 4;
      n_gx[j] = ld4(g_xout + base + c);
      n_xp[j] = ld4(xprime + base + c);
      n_gy[j] = *reinterpret_cast<const uint2*>(g_yln + base + c);
    }
    n_mean = stats[2 * row];
    n_rstd = stats[2 * row + 1];
  };
  int row = warp_global;
  if (row < n_rows) fetch(row);
  for (; row < n_rows; row += warps_total) {
    const long long base = (long long)row * D;
    float4 acc[NV], xh[NV], g[NV];
    const float mean = n_mean, rstd = n_rstd;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = (j * 32 + lane) * 4;
      acc[j] = n_gx[j];
      const float4 xv = n_xp[j];
      xh[j] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
      const float2 ga = unpack_bf16(n_gy[j].x), gb = unpack_bf16(n_gy[j].y);
      const float4 gy = make_float4(ga.x, ga.y, gb.x, gb.y);
      const float4 gm = ld4(gamma + c);
      ag[j].x += gy.x * xh[j].x; ag[j].y += gy.y * xh[j].y; ag[j].z += gy.z * xh[j].z; ag[j].w += gy.w * xh[j].w;
      ab[j].x += gy.x; ab[j].y += gy.y; ab[j].z += gy.z; ab[j].w += gy.w;
      g[j] = make_float4(gy.x * gm.x, gy.y * gm.y, gy.z * gm.z, gy.w * gm.w);
      s1 += (g[j].x + g[j].y) + (g[j].z + g[j].w);
      s2 += (g[j].x * xh[j].x + g[j].y * xh[j].y) + (g[j].z * xh[j].z + g[j].w * xh[j].w);
    }
    if (row + warps_total < n_rows) fetch(row + warps_total);  // in flight during the reductions and stores below
    if (g_ycast) {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float4 t = ldbf4(g_ycast + base + (j * 32 + lane) * 4);
        acc[j].x += t.x; acc[j].y += t.y; acc[j].z += t.z; acc[j].w += t.w;
      }
    }
    s1 = warp_sum(s1) * inv_d;
    s2 = warp_sum(s2) * inv_d;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = (j * 32 + lane) * 4;
      acc[j].x += rstd * (g[j].x - s1 - xh[j].x * s2);
      acc[j].y += rstd * (g[j].y - s1 - xh[j].y * s2);
      acc[j].z += rstd * (g[j].z - s1 - xh[j].z * s2);
      acc[j].w += rstd * (g[j].w - s1 - xh[j].w * s2);
      if (g_x) st4(g_x + base + c, acc[j]);
      if (g_branch) {
        float m[4];
        drop4(dc, (uint64_t)base + c, m);
        stbf4(g_branch + base + c,
              make_float4(alpha * m[0] * acc[j].x, alpha * m[1] * acc[j].y, alpha * m[2] * acc[j].z,
                          alpha * m[3] * acc[j].w));
      }
    }
  }
  if (dgamma) {
    flush_col_partials<NV>(ag, dgamma, red);
    flush_col_partials<NV>(ab, dbeta, red);
  }
}

// =================================================================================================
// ln_act:  h = dropout(gelu(LN(z)))  on bf16 rows  (feature_fusion / disc_* Sequential blocks)
// =================================================================================================
template <int NV>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
ln_act_fwd_kernel(const __nv_bfloat16* __restrict__ z, const float* __restrict__ gamma,
                  const float* __restrict__ beta, __nv_bfloat16* __restrict__ h,
                  float* __restrict__ stats, int n_rows, DropCfg dc_in) {
  const DropCfg dc = resolve_epoch(dc_in);
  constexpr int D = NV * 128;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const long long base = (long long)row * D;
  float4 v[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) v[j] = ldbf4(z + base + (j * 32 + lane) * 4);
  float mean, rstd;
  row_stats<NV>(v, mean, rstd);
  if (lane == 0) {
    stats[2 * row] = mean;
    stats[2 * row + 1] = rstd;
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (j * 32 + lane) * 4;
    const float4 g = ld4(gamma + c), b = ld4(beta + c);
    float m[4];
    drop4(dc, (uint64_t)base + c, m);
    float4 y;
    y.x = gelu_erf((v[j].x - mean) * rstd * g.x + b.x) * m[0];
    y.y = gelu_erf((v[j].y - mean) * rstd * g.y + b.y) * m[1];
    y.z = gelu_erf((v[j].z - mean) * rstd * g.z + b.z) * m[2];
    y.w = gelu_erf((v[j].w - mean) * rstd * g.w + b.w) * m[3];
    stbf4(h + base + c, y);
  }
}

template <int NV>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
ln_act_bwd_kernel(const __nv_bfloat16* __restrict__ g_h, const __nv_bfloat16* __restrict__ z,
                  const float* __restrict__ stats, const float* __restrict__ gamma,
                  const float* __restrict__ beta, __nv_bfloat16* __restrict__ g_z,
                  float* __restrict__ dgamma, float* __restrict__ dbeta, int n_rows, DropCfg dc_in) {
  const DropCfg dc = resolve_epoch(dc_in);
  constexpr int D = NV * 128;
  constexpr float inv_d = 1.0f / D;
  __shared__ float red[kWarpsPerBlock * D];
  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int warps_total = gridDim.x * kWarpsPerBlock;
  float4 ag[NV], ab[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) ag[j] = ab[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int row = warp_global; row < n_rows; row += warps_total) {
    const long long base = (long long)row * D;
    const float mean = stats[2 * row], rstd = stats[2 * row + 1];
    float4 xh[NV], g[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = (j * 32 + lane) * 4;
      const float4 zv = ldbf4(z + base + c);
      xh[j] = make_float4((zv.x - mean) * rstd, (zv.y - mean) * rstd, (zv.z - mean) * rstd,
                          (zv.w - mean) * rstd);
      const float4 gm = ld4(gamma + c), bt = ld4(beta + c);
      const float4 gh = ldbf4(g_h + base + c);
      float m[4];
      drop4(dc, (uint64_t)base + c, m);
      float4 du;
      du.x = gh.x * m[0] * gelu_erf_grad(xh[j].x * gm.x + bt.x);
      du.y = gh.y * m[1] * gelu_erf_grad(xh[j].y * gm.y + bt.y);
      du.z = gh.z * m[2] * gelu_erf_grad(xh[j].z * gm.z + bt.z);
      du.w = gh.w * m[3] * gelu_erf_grad(xh[j].w * gm.w + bt.w);
      ag[j].x += du.x * xh[j].x; ag[j].y += du.y * xh[j].y; ag[j].z += du.z * xh[j].z; ag[j].w += du.w * xh[j].w;
      ab[j].x += du.x; ab[j].y += du.y; ab[j].z += du.z; ab[j].w += du.w;
      g[j] = make_float4(du.x * gm.x, du.y * gm.y, du.z * gm.z, du.w * gm.w);
      s1 += (g[j].x + g[j].y) + (g[j].z + g[j].w);
      s2 += (g[j].x * xh[j].x + g[j].y * xh[j].y) + (g[j].z * xh[j].z + g[j].w * xh[j].w);
    }
    s1 = warp_sum(s1) * inv_d;
    s2 = warp_sum(s2) * inv_d;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = (j * 32 + lane) * 4;
      stbf4(g_z + base + c, make_float4(rstd * (g[j].x - s1 - xh[j].x * s2),
                                        rstd * (g[j].y - s1 - xh[j].y * s2),
                                        rstd * (g[j].z - s1 - xh[j].z * s2),
                                        rstd * (g[j].w - s1 - xh[j].w * s2)));
    }
  }
  flush_col_partials<NV>(ag, dgamma, red);
  flush_col_partials<NV>(ab, dbeta, red);
}

// =================================================================================================
// gelu_dropout (FFN activation): h = dropout(gelu(z));  dz = dh * mask * gelu'(z)
// =================================================================================================
__global__ void __launch_bounds__(256)
gelu_dropout_fwd_kernel(const __nv_bfloat16* __restrict__ z, __nv_bfloat16* __restrict__ h,
                        long long n8, DropCfg dc_in) {
  const DropCfg dc = resolve_epoch(dc_in);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    const uint4 u = *reinterpret_cast<const uint4*>(z + i * 8);
    const uint32_t in[4] = {u.x, u.y, u.z, u.w};
    uint32_t out[4];
    float mm[8];
    drop8(dc, (uint64_t)i * 8, mm);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack_bf16(in[k]);
      out[k] = pack_bf16(gelu_erf(f.x) * mm[2 * k], gelu_erf(f.y) * mm[2 * k + 1]);
    }
    *reinterpret_cast<uint4*>(h + i * 8) = make_uint4(out[0], out[1], out[2], out[3]);
  }
}

__global__ void __launch_bounds__(256)
gelu_dropout_bwd_kernel(const __nv_bfloat16* __restrict__ g_h, const __nv_bfloat16* __restrict__ z,
                        __nv_bfloat16* __restrict__ g_z, long long n8, DropCfg dc_in) {
  const DropCfg dc = resolve_epoch(dc_in);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    const uint4 uz = *reinterpret_cast<const uint4*>(z + i * 8);
    const uint4 ug = *reinterpret_cast<const uint4*>(g_h + i * 8);
    const uint32_t zi[4] = {uz.x, uz.y, uz.z, uz.w};
    const uint32_t gi[4] = {ug.x, ug.y, ug.z, ug.w};
    uint32_t out[4];
    float mm[8];
    drop8(dc, (uint64_t)i * 8, mm);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack_bf16(zi[k]);
      const float2 g = unpack_bf16(gi[k]);
      out[k] = pack_bf16(g.x * mm[2 * k] * gelu_erf_grad(f.x), g.y * mm[2 * k + 1] * gelu_erf_grad(f.y));
    }
    *reinterpret_cast<uint4*>(g_z + i * 8) = make_uint4(out[0], out[1], out[2], out[3]);
  }
}

// =================================================================================================
// colsum: out[n] += scale * sum_m X[m, n]   (bias gradients), bf16 in, fp32 accumulate.
// grid = (row splits, column slabs of 768); a warp streams whole rows of its slab (3 x 16 B per lane,
// 1.5 kB contiguous per warp access), 8 warps of a block own interleaved rows, partial sums meet in
// shared memory and leave as ONE atomicAdd per column per block.
// =================================================================================================
constexpr int kColsumSlab = 768;
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, long long ld, float* __restrict__ out, int M,
                   int N, int rows_per_block, float scale) {
  __shared__ float red[8][kColsumSlab];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.y * kColsumSlab;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(r0 + rows_per_block, M);
  float acc[3][8];
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[j][k] = 0.f;
  bool live[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) live[j] = c0 + (j * 32 + lane) * 8 < N;  // N % 8 == 0
#pragma unroll 4
  for (int r = r0 + warp; r < r1; r += 8) {
    const __nv_bfloat16* row = x + (long long)r * ld + c0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (live[j]) {
        const uint4 u = *reinterpret_cast<const uint4*>(row + (j * 32 + lane) * 8);
        const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
        acc[j][0] += a.x; acc[j][1] += a.y; acc[j][2] += b.x; acc[j][3] += b.y;
        acc[j][4] += c.x; acc[j][5] += c.y; acc[j][6] += d.x; acc[j][7] += d.y;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float* dst = &red[warp][(j * 32 + lane) * 8];
    *reinterpret_cast<float4*>(dst) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[j][4], acc[j][5], acc[j][6], acc[j][7]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < kColsumSlab; c += 256) {
    if (c0 + c < N) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][c];
      atomicAdd(out + c0 + c, s * scale);
    }
  }
}

// =================================================================================================
// small glue kernels
// =================================================================================================
// dst_bf16[r, col_off + c] = scale * src[r, c]   (src fp32 or bf16) — builds the feature_fusion input
// without torch.cat (model.py:450) and the bf16 GEMM views of fp32 tensors.
__global__ void __launch_bounds__(256)
cast_scale_kernel(const float* __restrict__ src_f32, const __nv_bfloat16* __restrict__ src_bf16,
                  long long ld_src, __nv_bfloat16* __restrict__ dst, long long rows, int cols,
                  long long ld_dst, int col_off, float scale) {
  const int c4n = cols / 4;
  const long long total = rows * c4n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c4n;
    const int c = (int)(i % c4n) * 4;
    float4 v = src_f32 ? ld4(src_f32 + r * ld_src + c) : ldbf4(src_bf16 + r * ld_src + c);
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    stbf4(dst + r * ld_dst + col_off + c, v);
  }
}

// out[b, :] = (1/S) * sum_s (x[b, s, :] (+ y_bf16[b, s, :]))   — sequence mean pooling (model.py:1193, 971)
__global__ void __launch_bounds__(256)
seq_mean_fwd_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ y,
                    float* __restrict__ out, int S, int D, int s_chunk) {
  const int b = blockIdx.y;
  const int s0 = blockIdx.x * s_chunk, s1 = min(s0 + s_chunk, S);
  const float inv = 1.0f / S;
  for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = s0; s < s1; ++s) {
      const long long o = ((long long)b * S + s) * D + c;
      if (x) { const float4 v = ld4(x + o); acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
      if (y) { const float4 v = ldbf4(y + o); acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
    }
    atomicAdd(out + (long long)b * D + c + 0, acc.x * inv);
    atomicAdd(out + (long long)b * D + c + 1, acc.y * inv);
    atomicAdd(out + (long long)b * D + c + 2, acc.z * inv);
    atomicAdd(out + (long long)b * D + c + 3, acc.w * inv);
  }
}

// g_x[b, s, :] (+)= g[b, :] / S   (fp32 accumulate or bf16 write) — backward of the pooling
__global__ void __launch_bounds__(256)
seq_mean_bwd_kernel(const float* __restrict__ g, float* __restrict__ gx_f32,
                    __nv_bfloat16* __restrict__ gy_bf16, long long rows, int S, int D) {
  const int c4n = D / 4;
  const long long total = rows * c4n;
  const float inv = 1.0f / S;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c4n;
    const int c = (int)(i % c4n) * 4;
    float4 v = ld4(g + (r / S) * D + c);
    v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
    if (gx_f32) st4(gx_f32 + r * D + c, v);
    if (gy_bf16) stbf4(gy_bf16 + r * D + c, v);
  }
}

inline int persistent_blocks(int n_rows) {
  int b = (n_rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const int cap = num_sms() * 4;
  return b < cap ? (b < 1 ? 1 : b) : cap;
}

}  // namespace
}  // namespace sct

using namespace sct;

#define DISPATCH_NV(d, ...)                                              \
  switch (d) {                                                           \
    case 384: { constexpr int NV = 3; __VA_ARGS__; } break;              \
    case 768: { constexpr int NV = 6; __VA_ARGS__; } break;              \
    case 1536: { constexpr int NV = 12; __VA_ARGS__; } break;            \
    default: SCT_CHECK(false, "unsupported row width d=%d (supported: 384, 768, 1536)", (int)(d)); \
  }

extern "C" {

int32_t sct_embed_ln_pe_fwd(const int64_t* ids, const float* table, const float* gamma,
                            const float* beta, const float* pe, float* out_f32, void* out_bf16,
                            float* stats, int64_t n_tok, int64_t seq_len, int64_t vocab, int64_t d,
                            float scale, float p_drop, uint64_t seed, uint64_t offset, const uint64_t* epoch, void* stream) {
  SCT_CHECK(ids && table && gamma && beta && pe && stats && (out_f32 || out_bf16), "null pointer");
  SCT_CHECK(n_tok > 0 && seq_len > 0, "empty input (n_tok=%lld seq_len=%lld)", (long long)n_tok, (long long)seq_len);
  const DropCfg dc = make_drop(p_drop, seed, offset, epoch);
  const int blocks = (int)((n_tok + kWarpsPerBlock - 1) / kWarpsPerBlock);
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_NV(d, (embed_ln_pe_fwd_kernel<NV><<<blocks, kWarpsPerBlock * 32, 0, st>>>(
                     ids, table, gamma, beta, pe, out_f32, (__nv_bfloat16*)out_bf16, stats, (int)n_tok,
                     (int)seq_len, (int)vocab, scale, dc)));
  SCT_LAUNCH_CHECK();
  return 0;
}

int32_t sct_embed_ln_pe_bwd(const float* g_f32, const void* g_bf16, const int64_t* ids,
                            const float* table, const float* gamma, const float* stats,
                            float* dtable, float* dgamma, float* dbeta, int64_t n_tok, int64_t vocab,
                            int64_t d, float scale, float p_drop, uint64_t seed, uint64_t offset, const uint64_t* epoch,
                            void* stream) {
  SCT_CHECK((g_f32 || g_bf16) && ids && table && gamma && stats && dtable && dgamma && dbeta, "null pointer");
  const DropCfg dc = make_drop(p_drop, seed, offset, epoch);
  const int blocks = persistent_blocks((int)n_tok);
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_NV(d, (embed_ln_pe_bwd_kernel<NV><<<blocks, kWarpsPerBlock * 32, 0, st>>>(
                     g_f32, (const __nv_bfloat16*)g_bf16, ids, table, gamma, stats, dtable, dgamma, dbeta,
                     (int)n_tok, (int)vocab, scale, dc)));
  SCT_LAUNCH_CHECK();
  return 0;
}

int32_t sct_add_dropout_ln_fwd(const float* x, const void* branch, float alpha, const float* gamma,
                               const float* beta, float* x_out, void* y_ln, void* y_cast,
                               float* stats, int64_t n_rows, int64_t d, float p_drop, uint64_t seed,
                               uint64_t offset, const uint64_t* epoch, void* stream) {
  SCT_CHECK(x || branch, "need x or branch");
  SCT_CHECK(!y_ln || (gamma && beta && stats), "LayerNorm output requested without gamma/beta/stats");
  SCT_CHECK(n_rows > 0, "empty input");
  const DropCfg dc = make_drop(p_drop, seed, offset, epoch);
  const int blocks = (int)((n_rows + kWarpsPerBlock - 1) / kWarpsPerBlock);
  cudaStream_t st = (cudaStream_t)stream;
  const int staged = env_int("SCT_LN_STAGED", 1);  // =0 selects the plain one-row-per-warp kernel (A/B timing)
  if (staged && d == 768 && n_rows >= 4096) {
    constexpr int NV = 6;
    constexpr int smem = kWarpsPerBlock * kLnSlots * (NV * 128 * 6) + kWarpsPerBlock * kLnSlots * 8;
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(add_dropout_ln_fwd_staged_kernel<NV>), smem)) return rc;
    int pb = 2 * num_sms();
    if (pb > blocks) pb = blocks;
    add_dropout_ln_fwd_staged_kernel<NV><<<pb, kWarpsPerBlock * 32, smem, st>>>(
        x, (const __nv_bfloat16*)branch, alpha, gamma, beta, x_out, (__nv_bfloat16*)y_ln, (__nv_bfloat16*)y_cast, stats,
        (int)n_rows, dc);
    SCT_LAUNCH_CHECK();
    return 0;
  }
  DISPATCH_NV(d, (add_dropout_ln_fwd_kernel<NV><<<blocks, kWarpsPerBlock * 32, 0, st>>>(
                     x, (const __nv_bfloat16*)branch, alpha, gamma, beta, x_out, (__nv_bfloat16*)y_ln,
                     (__nv_bfloat16*)y_cast, stats, (int)n_rows, dc)));
  SCT_LAUNCH_CHECK();
  return 0;
}

int32_t sct_add_dropout_ln_bwd(const float* g_xout, const void* g_yln, const void* g_ycast,
                               const float* xprime, const float* stats, const float* gamma,
                               float alpha, float* g_x, void* g_branch, float* dgamma, float* dbeta,
                               int64_t n_rows, int64_t d, float p_drop, uint64_t seed,
                               uint64_t offset, const uint64_t* epoch, void* stream) {
  SCT_CHECK(!g_yln || (xprime && stats && gamma && dgamma && dbeta), "LN backward needs x', stats, gamma, dgamma, dbeta");
  SCT_CHECK(g_x || g_branch, "no output requested");
  const DropCfg dc = make_drop(p_drop, seed, offset, epoch);
  const int blocks = persistent_blocks((int)n_rows);
  cudaStream_t st = (cudaStream_t)stream;
  if (g_yln && g_xout && d == 768 && n_rows >= 4096 && env_int("SCT_LN_BWD_PREFETCH", 1)) {  // =0: plain kernel (A/B)
    constexpr int NV = 6;
    int pb = num_sms();  // one block of 8 warps per SM is what the register file holds
    if (pb > blocks) pb = blocks;
    add_dropout_ln_bwd_pf_kernel<NV><<<pb, kWarpsPerBlock * 32, 0, st>>>(
        g_xout, (const __nv_bfloat16*)g_yln, (const __nv_bfloat16*)g_ycast, xprime, stats, gamma, alpha, g_x,
        (__nv_bfloat16*)g_branch, dgamma, dbeta, (int)n_rows, dc);
    SCT_LAUNCH_CHECK();
    return 0;
  }
  DISPATCH_NV(d, (add_dropout_ln_bwd_kernel<NV><<<blocks, kWarpsPerBlock * 32, 0, st>>>(
                     g_xout, (const __nv_bfloat16*)g_yln, (const __nv_bfloat16*)g_ycast, xprime, stats,
                     gamma, alpha, g_x, (__nv_bfloat16*)g_branch, dgamma, dbeta, (int)n_rows, dc)));
  SCT_LAUNCH_CHECK();
  return 0;
}

int32_t sct_ln_act_fwd(const void* z, const float* gamma, const float* beta, void* h, float* stats,
                       int64_t n_rows, int64_t d, float p_drop, uint64_t seed, uint64_t offset, const uint64_t* epoch,
                       void* stream) {
  SCT_CHECK(z && gamma && beta && h && stats, "null pointer");
  const DropCfg dc = make_drop(p_drop, seed, offset, epoch);
  const int blocks = (int)((n_rows + kWarpsPerBlock - 1) / kWarpsPerBlock);
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_NV(d, (ln_act_fwd_kernel<NV><<<blocks, kWarpsPerBlock * 32, 0, st>>>(
                     (const __nv_bfloat16*)z, gamma, beta, (__nv_bfloat16*)h, stats, (int)n_rows, dc)));
  SCT_LAUNCH_CHECK();
  return 0;
}

int32_t sct_ln_act_bwd(const void* g_h, const void* z, const float* stats, const float* gamma,
                       const float* beta, void* g_z, float* dgamma, float* dbeta, int64_t n_rows,
                       int64_t d, float p_drop, uint64_t seed, uint64_t offset, const uint64_t* epoch, void* stream) {
  SCT_CHECK(g_h && z && stats && gamma && beta && g_z && dgamma && dbeta, "null pointer");
  const DropCfg dc = make_drop(p_drop, seed, offset, epoch);
  const int blocks = persistent_blocks((int)n_rows);
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_NV(d, (ln_act_bwd_kernel<NV><<<blocks, kWarpsPerBlock * 32, 0, st>>>(
                     (const __nv_bfloat16*)g_h, (const __nv_bfloat16*)z, stats, gamma, beta,
                     (__nv_bfloat16*)g_z, dgamma, dbeta, (int)n_rows, dc)));
  SCT_LAUNCH_CHECK();
  return 0;
}

int32_t sct_gelu_dropout_fwd(const void* z, void* h, int64_t n, float p_drop, uint64_t seed,
                             uint64_t offset, const uint64_t* epoch, void* stream) {
  SCT_CHECK(z && h, "null pointer");
  SCT_CHECK(n % 8 == 0, "element count %lld not a multiple of 8", (long long)n);
  const DropCfg dc = make_drop(p_drop, seed, offset, epoch);
  const long long n8 = n / 8;
  long long blocks = (n8 + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  gelu_dropout_fwd_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)z, (__nv_bfloat16*)h, n8, dc);
  SCT_LAUNCH_CHECK();
  return 0;
}

int32_t sct_gelu_dropout_bwd(const void* g_h, const void* z, void* g_z, int64_t n, float p_drop,
                             uint64_t seed, uint64_t offset, const uint64_t* epoch, void* stream) {
  SCT_CHECK(g_h && z && g_z, "null pointer");
  SCT_CHECK(n % 8 == 0, "element count %lld not a multiple of 8", (long long)n);
  const DropCfg dc = make_drop(p_drop, seed, offset, epoch);
  const long long n8 = n / 8;
  long long blocks = (n8 + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  gelu_dropout_bwd_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)g_h, (const __nv_bfloat16*)z, (__nv_bfloat16*)g_z, n8, dc);
  SCT_LAUNCH_CHECK();
  return 0;
}

int32_t sct_colsum_bf16(const void* x, int64_t ld, float* out, int64_t M, int64_t N, float scale,
                        void* stream) {
  SCT_CHECK(x && out, "null pointer");
  SCT_CHECK(ld % 8 == 0 && N % 8 == 0, "colsum needs ld and N multiples of 8 (ld=%lld N=%lld)", (long long)ld, (long long)N);
  SCT_CHECK(M > 0 && N > 0, "empty input");
  const int slabs = (int)((N + kColsumSlab - 1) / kColsumSlab);
  int splits = (2 * num_sms() + slabs - 1) / slabs;
  int rpb = (int)((M + splits - 1) / splits);
  if (rpb < 64) rpb = 64;
  splits = (int)((M + rpb - 1) / rpb);
  dim3 grid((unsigned)splits, (unsigned)slabs);
  colsum_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ld, out, (int)M,
                                                            (int)N, rpb, scale);
  SCT_LAUNCH_CHECK();
  return 0;
}

int32_t sct_cast_scale(const float* src_f32, const void* src_bf16, int64_t ld_src, void* dst,
                       int64_t rows, int64_t cols, int64_t ld_dst, int64_t col_off, float scale,
                       void* stream) {
  SCT_CHECK((src_f32 != nullptr) != (src_bf16 != nullptr), "exactly one source");
  SCT_CHECK(cols % 4 == 0 && ld_dst % 4 == 0 && col_off % 4 == 0 && ld_src % 4 == 0 && ld_src >= cols,
            "cast_scale needs multiples of 4");
  const long long total = rows * (cols / 4);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  cast_scale_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
      src_f32, (const __nv_bfloat16*)src_bf16, ld_src, (__nv_bfloat16*)dst, rows, (int)cols, ld_dst,
      (int)col_off, scale);
  SCT_LAUNCH_CHECK();
  return 0;
}

int32_t sct_seq_mean_fwd(const float* x, const void* y_bf16, float* out, int64_t B, int64_t S,
                         int64_t d, void* stream) {
  SCT_CHECK((x || y_bf16) && out, "null pointer");
  SCT_CHECK(d % 4 == 0, "d must be a multiple of 4");
  SCT_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * B * d, (cudaStream_t)stream));
  int chunks = (int)((num_sms() * 4 + B - 1) / B);
  if (chunks < 1) chunks = 1;
  int s_chunk = (int)((S + chunks - 1) / chunks);
  if (s_chunk < 8) s_chunk = 8;
  chunks = (int)((S + s_chunk - 1) / s_chunk);
  dim3 grid(chunks, (unsigned)B);
  seq_mean_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, (const __nv_bfloat16*)y_bf16, out, (int)S,
                                                             (int)d, s_chunk);
  SCT_LAUNCH_CHECK();
  return 0;
}

int32_t sct_seq_mean_bwd(const float* g, float* gx_f32, void* gy_bf16, int64_t B, int64_t S, int64_t d,
                         void* stream) {
  SCT_CHECK(g && (gx_f32 || gy_bf16), "null pointer");
  const long long rows = B * S;
  const long long total = rows * (d / 4);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  seq_mean_bwd_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(g, gx_f32, (__nv_bfloat16*)gy_bf16, rows,
                                                                    (int)S, (int)d);
  SCT_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
