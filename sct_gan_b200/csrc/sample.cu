// sct_gan_b200 — sampling tail of the generation loop.
//
// Replaces, per decode step and for all B rows at once (reference model.py:892-918):
//   logits = logits / temperature;  logits[:, 59] *= 2 after ids 2000..2002 (the one live effect of
//   _apply_syntax_constraints, model.py:975-1060);  top-k filter;  nucleus (top-p) filter on the sorted survivors;
//   softmax;  multinomial draw            — or argmax when greedy —
// which is ~12 PyTorch launches over [B, 50265] fp32 tensors (divide, clone, where, topk, softmax, cumsum, compare,
// shift, masked_fill, softmax, multinomial, gather) with ONE launch that reads the bf16 logits row once.
//
// One block per row.  A thread keeps a CONTIGUOUS slice of the row (13 x 8 bf16) in registers, so every later step is
// register arithmetic:
//   1. the top_k-th largest logit by bisection over the 16-bit ordered-key space of bf16 (16 block-wide counts with
//      packed bf16x2 comparisons — a shared-memory histogram would serialise on the handful of exponent bytes that
//      hold all logits);
//   2. survivors = everything above that value plus, among the elements EQUAL to it, the lowest indices (a block scan
//      over the per-thread tie counts keeps index order; bf16 logits tie often) — at most top_k (<= 64) entries;
//   3. rank sort of the survivors (value descending, index ascending: the order torch.sort / argmax give);
//   4. one thread: softmax over the survivors in fp32, the reference's shifted cumulative-probability cut, inverse-CDF
//      draw with a counter-based uniform (seed, offset, device-resident epoch, row).
// Temperature and the x2 tweak are monotone, so selection runs on the raw bf16 values (the tweak doubles the bf16
// element exactly); the fp32 arithmetic of steps 3-4 is the reference's: (l / T), then softmax.
#include <cuda_bf16.h>

#include "common.cuh"
#include "../../include/sct_b200.h"

namespace sct {
namespace {

constexpr int kSampleThreads = 512;
constexpr int kVecPerThread = 13;                       // 13 x 8 bf16 per thread: rows up to 53,248 entries
constexpr int kMaxTopK = 64;
constexpr int kMaxV = kSampleThreads * kVecPerThread * 8;

// bf16 bit pattern <-> 16-bit key whose unsigned order is the numeric order
__device__ __forceinline__ uint32_t key_of(uint32_t bits) { return (bits & 0x8000u) ? (~bits & 0xFFFFu) : (bits | 0x8000u); }
__device__ __forceinline__ uint32_t bits_of(uint32_t key) { return (key & 0x8000u) ? (key & 0x7FFFu) : (~key & 0xFFFFu); }

__device__ __forceinline__ __nv_bfloat162 as_bf162(uint32_t w) {
  __nv_bfloat162 v;
  *reinterpret_cast<uint32_t*>(&v) = w;
  return v;
}

__device__ __forceinline__ int block_sum(int v, int* red) {
  v = __reduce_add_sync(0xffffffffu, v);
  __syncthreads();  // red may still be read from the previous call
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  int s = 0;
#pragma unroll
  for (int w = 0; w < kSampleThreads / 32; ++w) s += red[w];
  return s;
}

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__global__ void __launch_bounds__(kSampleThreads)
sample_rows_kernel(const __nv_bfloat16* __restrict__ logits, long long ld, int V, float temperature, int top_k,
                   float top_p, int greedy, const long long* __restrict__ prev, unsigned long long seed,
                   unsigned long long offset, const unsigned long long* __restrict__ epoch,
                   long long* __restrict__ next_out) {
  __shared__ int red[kSampleThreads / 32];
  __shared__ int scan[kSampleThreads / 32];
  __shared__ int n_list;
  __shared__ uint32_t l_key[kMaxTopK], l_idx[kMaxTopK];
  __shared__ float s_val[kMaxTopK];
  __shared__ uint32_t s_idx[kMaxTopK];
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const __nv_bfloat16* x = logits + (long long)row * ld;
  const bool tweak = prev != nullptr && V > 59 && prev[row] >= 2000 && prev[row] <= 2002;
  const int e0 = tid * (kVecPerThread * 8);  // first element of this thread's slice

  // ---- the slice, as packed bf16 pairs; entries past V become -inf
  uint32_t w[kVecPerThread * 4];
#pragma unroll
  for (int v = 0; v < kVecPerThread; ++v) {
    const int e = e0 + v * 8;
    uint4 u = make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);
    if (e + 8 <= V) {
      u = *reinterpret_cast<const uint4*>(x + e);  // rows are 16-byte aligned (ld % 8 == 0)
    } else if (e < V) {
      uint32_t h[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) h[i] = e + i < V ? (uint32_t)__bfloat16_as_ushort(x[e + i]) : 0xFF80u;
      u = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
    }
    w[4 * v] = u.x, w[4 * v + 1] = u.y, w[4 * v + 2] = u.z, w[4 * v + 3] = u.w;
  }
  if (tweak && e0 <= 59 && 59 < e0 + kVecPerThread * 8) {  // logits[:, 59] *= 2 (exact in bf16; model.py:1049-1056)
    const int j = 59 - e0;
#pragma unroll
    for (int q = 0; q < kVecPerThread * 4; ++q) {
      if (q == (j >> 1)) {
        const uint32_t h = (j & 1) ? (w[q] >> 16) : (w[q] & 0xFFFFu);
        const float d = 2.0f * __uint_as_float(h << 16);
        const uint32_t nb = (uint32_t)__bfloat16_as_ushort(__float2bfloat16(d));
        w[q] = (j & 1) ? ((w[q] & 0xFFFFu) | (nb << 16)) : ((w[q] & 0xFFFF0000u) | nb);
      }
    }
  }
  const int k = min(min(greedy ? 1 : top_k, kMaxTopK), V);

  // ---- 1. largest bf16 value t with #{x >= t} >= k: bisection over the ordered keys
  auto count_ge = [&](uint32_t key) {
    const uint32_t b = bits_of(key);
    const __nv_bfloat162 t2 = as_bf162(b | (b << 16));
    int c = 0;
#pragma unroll
    for (int q = 0; q < kVecPerThread * 4; ++q) c += __popc(__hge2_mask(as_bf162(w[q]), t2));
    return block_sum(c, red) >> 4;  // 16 mask bits per element
  };
  uint32_t lo = key_of(0xFF80u), hi = key_of(0x7F80u) + 1;  // [-inf, +inf]: count_ge(lo) = V >= k
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (count_ge(mid) >= k) lo = mid;
    else hi = mid;
  }
  const uint32_t tb = bits_of(lo);
  const float tval = __uint_as_float(tb << 16);
  // ---- 2. survivors: above the threshold (fewer than k), then ties in index order
  if (tid == 0) n_list = 0;
  __syncthreads();
  int n_eq = 0;
#pragma unroll
  for (int q = 0; q < kVecPerThread * 4; ++q) {
#pragma unroll
    for (int hsel = 0; hsel < 2; ++hsel) {
      const uint32_t h = hsel ? (w[q] >> 16) : (w[q] & 0xFFFFu);
      const float f = __uint_as_float(h << 16);
      if (f > tval) {
        const int s = atomicAdd(&n_list, 1);
        if (s < kMaxTopK) l_key[s] = key_of(h), l_idx[s] = (uint32_t)(e0 + 2 * q + hsel);
      } else if (f == tval) {
        ++n_eq;
      }
    }
  }
  // exclusive block scan of the tie counts (threads own contiguous, increasing index ranges)
  int incl = n_eq;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) scan[warp] = incl;
  __syncthreads();  // also orders the n_list atomics before the read below
  int before = incl - n_eq;
  for (int i = 0; i < warp; ++i) before += scan[i];
  const int n_gt = n_list;
  const int need = k - n_gt;  // >= 1 by construction of the threshold
  if (n_eq > 0 && before < need) {
    int r = before;
#pragma unroll
    for (int q = 0; q < kVecPerThread * 4; ++q) {
#pragma unroll
      for (int hsel = 0; hsel < 2; ++hsel) {
        const uint32_t h = hsel ? (w[q] >> 16) : (w[q] & 0xFFFFu);
        if (__uint_as_float(h << 16) == tval) {
          if (r < need) l_key[n_gt + r] = key_of(h), l_idx[n_gt + r] = (uint32_t)(e0 + 2 * q + hsel);
          ++r;
        }
      }
    }
  }
  __syncthreads();
  // ---- 3. rank sort: value descending, index ascending
  if (tid < k) {
    const uint32_t mk = l_key[tid], mi = l_idx[tid];
    int r = 0;
    for (int i = 0; i < k; ++i) r += (l_key[i] > mk) || (l_key[i] == mk && l_idx[i] < mi);
    s_val[r] = __uint_as_float(bits_of(mk) << 16);
    s_idx[r] = mi;
  }
  __syncthreads();
  // ---- 4. softmax over the survivors, nucleus cut, draw
  if (tid == 0) {
    int pick = 0;
    if (!greedy) {
      float e[kMaxTopK];
      const float m = __fdiv_rn(s_val[0], temperature);
      float Z = 0.f;
      for (int j = 0; j < k; ++j) {
        e[j] = expf(__fdiv_rn(s_val[j], temperature) - m);
        Z += e[j];
      }
      // reference: remove token j when the cumulative probability up to j - 1 already exceeds top_p (token 0 stays)
      int keep = 1;
      float cum = e[0] / Z, Zk = e[0];
      while (keep < k && !(cum > top_p)) {
        Zk += e[keep];
        cum += e[keep] / Z;
        ++keep;
      }
      const unsigned long long ep = epoch ? *epoch : 0ull;
      const uint64_t h = mix64(mix64(seed ^ (offset * 0x9E3779B97F4A7C15ull)) ^ mix64(ep * 0xD1B54A32D192ED03ull + (uint64_t)row));
      const float u = (float)(h >> 40) * (1.0f / 16777216.0f);  // [0, 1)
      const float target = u * Zk;
      float acc = 0.f;
      pick = keep - 1;
      for (int j = 0; j < keep; ++j) {
        acc += e[j];
        if (acc > target) {
          pick = j;
          break;
        }
      }
    }
    next_out[row] = (long long)s_idx[pick];
  }
}

}  // namespace
}  // namespace sct

extern "C" int32_t sct_sample_rows(const void* logits, int64_t ld, int64_t n_rows, int64_t V, float temperature,
                                   int32_t top_k, float top_p, int32_t greedy, const int64_t* prev_tokens,
                                   uint64_t seed, uint64_t offset, const uint64_t* epoch, int64_t* next_tokens,
                                   void* stream) {
  using namespace sct;
  SCT_CHECK(logits && next_tokens, "null pointer");
  SCT_CHECK(n_rows > 0 && V > 0, "empty input");
  SCT_CHECK(V <= kMaxV, "vocabulary of %lld entries exceeds the %d a block holds in registers", (long long)V, kMaxV);
  SCT_CHECK(ld % 8 == 0 && ld >= V, "logits pitch must be a multiple of 8 and >= V");
  SCT_CHECK((reinterpret_cast<uintptr_t>(logits) & 15) == 0, "logits must be 16-byte aligned");
  SCT_CHECK(greedy || (top_k >= 1 && top_k <= kMaxTopK), "top_k must be in [1, %d] (got %d)", kMaxTopK, (int)top_k);
  SCT_CHECK(greedy || (temperature > 0.f && top_p > 0.f), "temperature and top_p must be positive");
  sample_rows_kernel<<<(unsigned)n_rows, kSampleThreads, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)logits, ld, (int)V, temperature, top_k, top_p, greedy,
      reinterpret_cast<const long long*>(prev_tokens), seed, offset,
      reinterpret_cast<const unsigned long long*>(epoch), reinterpret_cast<long long*>(next_tokens));
  SCT_LAUNCH_CHECK();
  return 0;
}
