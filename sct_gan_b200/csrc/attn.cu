// sct_b200 — K3: fused multi-head attention (forward + backward) on tcgen05 / TMEM, head_dim = 96.
//
// Replaces F.scaled_dot_product_attention / the bmm+softmax+dropout+bmm path of nn.MultiheadAttention
// for every attention site of the reference (SCT-GAN/model.py:56-77 encoder/decoder layers, :209-222
// ast_attention / cross_attention, :241-246 disc_path_attention; torch nn/functional.py:6244-6691):
// softmax(q k^T / sqrt(dh) + key_padding(-inf) + causal(-inf)) -> dropout -> @ v, never materialising
// the [B,H,Lq,Lk] probabilities.  The averaged attention weights the reference returns from its
// top-level MHAs are discarded by every caller (model.py:433,443,1186), so they are not produced.
//
// Data layout: Q/K/V/O and their gradients stay in the projection GEMMs' row-major layouts
// ([B*L, ld] with head h in columns [96h, 96h+96)); tiles are fetched by 3-D TMA ([B, L, cols] so rows
// past L zero-fill) as three [rows x 32 col] blocks with the 64-byte swizzle.  The same smem tile is
// consumed K-major (rows = M/N, e.g. Q and K in S = Q K^T) or MN-major (rows = K, e.g. V in O = P V)
// purely by choice of UMMA descriptor, so nothing is ever transposed.
//
// TMEM accumulator rows map 1:1 to threads (tcgen05.ld 32x32b), so the online softmax needs no
// cross-thread reduction.  Masks are applied from the [B,Lk] key-padding bytes and the causal
// predicate; dropout is regenerated from (seed, offset, element index).
//
// Kernels in this file:
//   attn_fwd_kernel         forward, one CTA per (128-query tile, head, batch), two per SM; P returns to TMEM as the A
//                           operand of P V; O leaves through a TMA store
//   attn_decode_kernel      forward for Lq = 1 (generation): SIMT stream over the K/V cache
//   attn_bwd_dvec_kernel    D = rowsum(dO * O)
//   attn_bwd_dkdv_kernel    dK, dV per key tile (two MMA issuers, P^T in TMEM); also writes dS^T to a workspace
//   attn_bwd_dq_ds_kernel   dQ = scale * dS K as a TMA -> tcgen05 stream over that workspace
//   attn_bwd_dq_kernel      dQ by recomputation (workspace == NULL)
#include <math.h>
#include <stdlib.h>

#include "../../include/sct_b200.h"
#include "common.cuh"

namespace sct {
namespace {

constexpr int DH = 96;            // head dim
constexpr int TILE = 128;         // q-tile and kv-tile rows
constexpr int BLK = TILE * 64;    // bytes of one [128 x 32 bf16] swizzle-64 column block
constexpr int QKV_BYTES = 3 * BLK;          // one [128 x 96] operand tile
constexpr float kLog2e = 1.4426950408889634f;

struct AttnParams {
  int B, H, Lq, Lk;
  int ldo;               // O / dO row pitch (elements)
  int ldq_out, ldkv_out; // pitches of dQ and dK/dV outputs
  int causal;
  float scale_log2;      // scale * log2(e)
  float scale;
  const uint8_t* kpm;    // [B, Lk], 1 = ignore key; nullable
  float* lse2;           // [B, H, Lq]  log2-domain logsumexp of the scaled scores
  float* dvec;           // [B, H, Lq]  D = rowsum(dO * O)
  __nv_bfloat16* o;      // forward output
  __nv_bfloat16 *dq, *dk, *dv;
  const unsigned long long* epoch;  // device-resident dropout epoch (nullable), folded into the key at run time
  uint32_t key0, key1;   // dropout stream key, mixed from (seed, offset) on the host
  uint32_t tmask[16];    // bit i of the kDropBits-bit drop threshold, spread to a full word
  uint32_t thresh16;     // the threshold (kDropBits bits); 0 = dropout off
  float inv_keep;
  int write_ds;          // dK/dV kernel: also store the dS^T tiles to the workspace (for the streaming dQ kernel)
};

// ---- attention-probability dropout -------------------------------------------------------------
// kDropBits: resolution of the drop probability (p is rounded to a multiple of 2^-8: 0.3 -> 77/256 = 0.3008; the scale
// uses the realised keep probability, so the expectation is exact).  Every bit is one round of the bit-sliced
// generator, i.e. 4 integer instructions per 32 probabilities, in the forward AND in both backward kernels — and
// integer / logic instructions run at half rate on sm_100 (tools/micro/alu_rates.cu: LOP3 / PRMT / IMAD 64 per clock
// per SM against 128 for FFMA), so every round costs as much pipe time as 8 multiply-adds per 32 scores.
constexpr int kDropBits = 8;
// keep(b,h,q,k) <=> u(b,h,q,k) >= thresh.  The kDropBits-bit uniforms of 32 consecutive keys of one query
// row are generated bit-sliced: 16 cheap xorshift-multiply words off one strong hash of (stream key,
// b*H+h, q, k/32); a 16-step bitwise comparator (one LOP3 per word) then yields the 32 keep bits at
// once, ~2 integer instructions per element instead of one full hash each.  Forward and the dQ kernel
// own query rows and use the word directly; the dK/dV kernel owns key rows and transposes 32x32 bit
// tiles inside the warp, so all three regenerate the identical mask.
__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}
struct DropKey {
  uint32_t k0, k1;
};
__device__ __forceinline__ DropKey drop_key(const AttnParams& p) {
  DropKey k = {p.key0, p.key1};
  if (p.thresh16 != 0 && p.epoch != nullptr) {
    const unsigned long long e = *p.epoch;
    k.k0 ^= fmix32((uint32_t)e * 0x9E3779B1u + 0x68E31DA4u);
    k.k1 += fmix32((uint32_t)(e >> 32) ^ 0xB5297A4Du) + (uint32_t)e;
  }
  return k;
}
__device__ __forceinline__ uint32_t keep_word(const AttnParams& p, const DropKey& dk, uint32_t bh, uint32_t q,
                                              uint32_t kb) {
  if (p.thresh16 == 0) return 0xFFFFFFFFu;
  uint32_t x = fmix32(((bh * (uint32_t)p.Lq + q) * 0x9E3779B1u) ^ dk.k0);
  x = fmix32(x ^ (kb * 0x85EBCA77u) ^ dk.k1);
  uint32_t lt = 0u;  // lt bit = 1 <=> u < thresh (LSB-first ripple comparison)
#pragma unroll
  for (int i = 0; i < kDropBits; ++i) {
    x = (x ^ (x >> 15)) * 0x2C1B3C6Du;
    const uint32_t nw = ~x, tm = p.tmask[i];
    lt = (nw & lt) | (tm & (nw | lt));
  }
  return ~lt;
}
// 32x32 bit-matrix transpose across the warp: in: lane L holds row L; out: lane L holds column L.
__device__ __forceinline__ uint32_t warp_bit_transpose(uint32_t x, int lane) {
#pragma unroll
  for (int sft = 16; sft > 0; sft >>= 1) {
    const uint32_t m = sft == 16 ? 0x0000FFFFu : sft == 8 ? 0x00FF00FFu : sft == 4 ? 0x0F0F0F0Fu
                       : sft == 2 ? 0x33333333u : 0x55555555u;
    const uint32_t y = __shfl_xor_sync(0xffffffffu, x, sft);
    x = (lane & sft) ? (((y >> sft) & m) | (x & ~m)) : ((x & m) | ((y & m) << sft));
  }
  return x;
}
// last unmasked key + 1 of batch row b (keys past it are all masked: their tiles are skipped)
__device__ __forceinline__ int kv_extent(const AttnParams& p, int b, int* smem_slot) {
  if (p.kpm == nullptr) return p.Lk;
  if (threadIdx.x == 0) *smem_slot = 0;
  __syncthreads();
  int loc = 0;
  for (int i = threadIdx.x; i < p.Lk; i += blockDim.x)
    if (p.kpm[(long long)b * p.Lk + i] == 0) loc = i + 1;
  loc = max(loc, __shfl_xor_sync(0xffffffffu, loc, 16));
  loc = max(loc, __shfl_xor_sync(0xffffffffu, loc, 8));
  loc = max(loc, __shfl_xor_sync(0xffffffffu, loc, 4));
  loc = max(loc, __shfl_xor_sync(0xffffffffu, loc, 2));
  loc = max(loc, __shfl_xor_sync(0xffffffffu, loc, 1));
  if ((threadIdx.x & 31) == 0) atomicMax(smem_slot, loc);
  __syncthreads();
  return *smem_slot;
}

// Descriptor low words (see common.cuh: the high word is a compile-time constant, the single issuing thread only
// adds immediates).  kHi64 / kHi128: 8-row groups 512 B / 1024 B apart, swizzle 64 / 128.
constexpr uint32_t kHi64 = umma_desc_hi(512, UMMA_SW64);
constexpr uint32_t kHi128 = umma_desc_hi(1024, UMMA_SW128);
// K-major swizzle-64 operand tile ([rows x 96], 3 blocks), k16 step k (0..5)
__device__ __forceinline__ uint32_t lo_k64(uint32_t tile, int k) {
  return umma_desc_lo(tile + (k >> 1) * BLK + (k & 1) * 32, 16);
}
// MN-major view of the same tile (rows = contraction index), k16 step k (0..7): N = 96 spans the three column
// blocks (LBO = BLK), 8-row groups are 512 B apart (SBO), 16 rows per step = 1024 B.
__device__ __forceinline__ uint32_t lo_mn64(uint32_t tile, int k) { return umma_desc_lo(tile + k * 1024, BLK); }
// The MMA-issuing thread is a serial chain of dependent integer instructions between tcgen05.mma's (measured: ~18
// SASS instructions and ~90 clk per MMA when every descriptor is rebuilt with shift / mask / or, against 54 clk for the
// bare instruction, tools/micro/mma_issue.cu), so the loops below build each tile's low word ONCE and step through
// it with immediate adds (byte offset >> 4; operand tiles never cross the 14-bit address field).
__device__ __forceinline__ uint32_t base_k64(uint32_t tile) { return umma_desc_lo(tile, 16); }
__device__ __forceinline__ uint32_t base_mn64(uint32_t tile) { return umma_desc_lo(tile, BLK); }
__device__ __forceinline__ uint32_t step_k64(uint32_t base, int k) { return base + (((k >> 1) * BLK + (k & 1) * 32) >> 4); }
__device__ __forceinline__ uint32_t step_mn64(uint32_t base, int k) { return base + ((k * 1024) >> 4); }

// write 32 consecutive bf16 (columns c0..c0+31, c0 % 32 == 0) of row r into a swizzle-128 [128x128] tile
__device__ __forceinline__ void store_row32_sw128(uint32_t tile, int r, int c0, const float (&v)[32]) {
  const uint32_t rowbase = tile + (c0 >> 6) * (TILE * 128) + r * 128;
  const int chunk0 = (c0 & 63) >> 3;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t dst = rowbase + ((((chunk0 + q) ^ (r & 7)) & 7) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst),
                 "r"(pack_bf16(v[8 * q], v[8 * q + 1])), "r"(pack_bf16(v[8 * q + 2], v[8 * q + 3])),
                 "r"(pack_bf16(v[8 * q + 4], v[8 * q + 5])), "r"(pack_bf16(v[8 * q + 6], v[8 * q + 7]))
                 : "memory");
  }
}

// write 16 consecutive bf16 (columns c0..c0+15, c0 % 16 == 0) of row r into a [128 x 64] swizzle-128 block
__device__ __forceinline__ void store_row16_sw128(uint32_t blk, int r, int c0, const float (&v)[16]) {
  const uint32_t rowbase = blk + r * 128;
  const int chunk0 = c0 >> 3;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const uint32_t dst = rowbase + ((((chunk0 + q) ^ (r & 7)) & 7) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst),
                 "r"(pack_bf16(v[8 * q], v[8 * q + 1])), "r"(pack_bf16(v[8 * q + 2], v[8 * q + 3])),
                 "r"(pack_bf16(v[8 * q + 4], v[8 * q + 5])), "r"(pack_bf16(v[8 * q + 6], v[8 * q + 7]))
                 : "memory");
  }
}

// epilogues: 32 fp32 accumulator columns of row r (scaled) -> bf16 into column block c of a swizzle-64 [128 x 96]
// tile (the layout the 3-D tensor maps of Q / K / V / O use), from where ONE thread stores the tile with TMA: a
// thread-per-row st.global touches 32 different 4.6 kB-strided rows per instruction and took ~15 % of the dK/dV
// kernel's time (ncu source view).
__device__ __forceinline__ void stage_chunk_sw64(uint32_t tile, int r, int c, const uint32_t (&rr)[32], float sc) {
  const uint32_t rowbase = tile + c * BLK + r * 64;
#pragma unroll
  for (int qd = 0; qd < 4; ++qd) {
    const uint32_t dst = rowbase + (((qd ^ (r >> 1)) & 3) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst),
                 "r"(pack_bf16(__uint_as_float(rr[8 * qd]) * sc, __uint_as_float(rr[8 * qd + 1]) * sc)),
                 "r"(pack_bf16(__uint_as_float(rr[8 * qd + 2]) * sc, __uint_as_float(rr[8 * qd + 3]) * sc)),
                 "r"(pack_bf16(__uint_as_float(rr[8 * qd + 4]) * sc, __uint_as_float(rr[8 * qd + 5]) * sc)),
                 "r"(pack_bf16(__uint_as_float(rr[8 * qd + 6]) * sc, __uint_as_float(rr[8 * qd + 7]) * sc))
                 : "memory");
  }
}
__device__ __forceinline__ void store_tile(const CUtensorMap* m, uint32_t src, int col0, int row0, int b) {
#pragma unroll
  for (int c = 0; c < 3; ++c) tma_store_3d(m, src + c * BLK, col0 + 32 * c, row0, b);
}

__device__ __forceinline__ void load_tile(const CUtensorMap* m, uint32_t bar, uint32_t dst, int col0,
                                          int row0, int b) {
#pragma unroll
  for (int c = 0; c < 3; ++c) tma_load_3d(m, bar, dst + c * BLK, col0 + 32 * c, row0, b);
}

// =================================================================================================
// forward: one CTA (256 threads) per (q-tile, head, batch); two CTAs co-reside per SM so one CTA's
// softmax overlaps the other's MMAs.  Thread (r, hf) owns columns [64hf, 64hf+64) of query row r: the
// S tile is read from TMEM exactly once (64 registers), the two halves of a row exchange their partial
// max through shared memory.  The running max is only raised when it grows by more than 2^8 (the
// probabilities then stay below 2^8, harmless in fp32/bf16), so the O accumulator in TMEM is almost
// never rescaled after the first tile.  Key tiles past the last unmasked key are skipped.
// =================================================================================================
constexpr int FWD_SMEM = 1024 + 3 * QKV_BYTES + 4096;

__global__ void __launch_bounds__(256, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQ = base, sK = sQ + QKV_BYTES, sV = sK + QKV_BYTES;
  const uint32_t aux = sV + QKV_BYTES;
  float* bias_s = reinterpret_cast<float*>(gen + (aux - base));  // [128]
  float* mx_s = bias_s + 128;                                     // [2][128]
  float* l_s = mx_s + 256;                                        // [2][128]
  const uint32_t bar_q = aux + 3072, bar_k = aux + 3080, bar_v = aux + 3088, bar_mma = aux + 3096;
  const uint32_t tmem_ptr_addr = aux + 3104;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(gen + (tmem_ptr_addr - base));
  int* ext_slot = reinterpret_cast<int*>(gen + (aux + 3112 - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, hf = warp >> 2;
  const int nq_tiles = gridDim.x;
  const int qt = nq_tiles - 1 - blockIdx.x;  // heavy (late) causal tiles first
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * TILE;

  if (tid == 0) {
    mbar_init(bar_q, 1);
    mbar_init(bar_k, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr_addr, 256);
  const int kv_end = kv_extent(p, b, ext_slot);  // contains __syncthreads when a mask is given
  int nkv = (kv_end + TILE - 1) / TILE;
  if (p.causal) nkv = min(nkv, qt + 1);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_gen;
  const uint32_t tS = tmem, tO = tmem + 128;
  const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;

  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, false, false);
  constexpr uint32_t idesc_o = umma_idesc_bf16(128, DH, false, true);
  const uint32_t bQ = base_k64(sQ), bK = base_k64(sK), bVm = base_mn64(sV);  // descriptor low words, built once

  if (tid == 0 && nkv > 0) {
    mbar_expect_tx(bar_q, QKV_BYTES);
    load_tile(&tmQ, bar_q, sQ, h * DH, q0, b);
    mbar_expect_tx(bar_k, QKV_BYTES);
    load_tile(&tmK, bar_k, sK, h * DH, 0, b);
    mbar_expect_tx(bar_v, QKV_BYTES);
    load_tile(&tmV, bar_v, sV, h * DH, 0, b);
    mbar_wait(bar_q, 0);
    mbar_wait(bar_k, 0);
    tc_fence_after();
#pragma unroll
    for (int k = 0; k < 6; ++k) tc_mma_bf16_lh(tS, step_k64(bQ, k), kHi64, step_k64(bK, k), kHi64, idesc_s, k > 0);
    tc_commit(bar_mma);
  }

  const int r = quad * 32 + lane;  // local q row
  const int q = q0 + r;
  const int c_base = hf * 64;
  const uint32_t bh = (uint32_t)(b * p.H + h);
  const DropKey dkey = drop_key(p);
  float m_run = -INFINITY, l_run = 0.f;

  for (int j = 0; j < nkv; ++j) {
    const int kv0 = j * TILE;
    bool masked = false;
    if (tid < TILE) {
      const int kv = kv0 + tid;
      masked = kv >= p.Lk;
      if (!masked && p.kpm) masked = p.kpm[(long long)b * p.Lk + kv] != 0;
      bias_s[tid] = masked ? -INFINITY : 0.f;
    }
    const int any_mask = __syncthreads_or(masked ? 1 : 0);
    mbar_wait(bar_mma, j & 1);  // S_j ready; P V_{j-1} retired
    tc_fence_after();
    if (tid == 0) {
      if (j + 1 < nkv) {
        mbar_expect_tx(bar_k, QKV_BYTES);
        load_tile(&tmK, bar_k, sK, h * DH, kv0 + TILE, b);
      }
      if (j > 0) {
        mbar_expect_tx(bar_v, QKV_BYTES);
        load_tile(&tmV, bar_v, sV, h * DH, kv0, b);
      }
    }
    __syncwarp();
    const bool diag = p.causal && (j == qt);
    uint32_t sr[64];
    {
      uint32_t (&lo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sr[0]);
      uint32_t (&hi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sr[32]);
      tmem_ld32(tS + lane_sel + c_base, lo);
      tmem_ld32(tS + lane_sel + c_base + 32, hi);
      tmem_ld_wait();
    }
    if (any_mask) {
#pragma unroll
      for (int i = 0; i < 64; ++i) sr[i] = __float_as_uint(__uint_as_float(sr[i]) + bias_s[c_base + i]);
    }
    if (diag) {
#pragma unroll
      for (int i = 0; i < 64; ++i)
        if (c_base + i > r) sr[i] = __float_as_uint(-INFINITY);
    }
    float mx = __uint_as_float(sr[0]);
#pragma unroll
    for (int i = 1; i < 64; ++i) mx = fmaxf(mx, __uint_as_float(sr[i]));
    mx_s[hf * 128 + r] = mx;
    named_bar_sync(1, 256);
    mx = fmaxf(mx, mx_s[(hf ^ 1) * 128 + r]);
    const float m_new = mx * p.scale_log2;  // scale > 0
    float m_next = m_run;
    if (m_run == -INFINITY || m_new > m_run + 8.0f) m_next = fmaxf(m_run, m_new);
    const float alpha = (m_run == -INFINITY) ? 0.f : exp2f(m_run - m_next);
    const float m_eff = (m_next == -INFINITY) ? 0.f : m_next;
    // packed fp32 arithmetic (FFMA2 / FADD2: two scores per instruction) for the scale-and-shift and the row sum
    const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(-m_eff, -m_eff);
    float2 rs2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int kb = 0; kb < 2; ++kb) {
      const uint32_t kw = keep_word(p, dkey, bh, (uint32_t)q, (uint32_t)((kv0 + c_base) >> 5) + kb);
      float pv[32];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float2 x = __ffma2_rn(make_float2(__uint_as_float(sr[kb * 32 + i]), __uint_as_float(sr[kb * 32 + i + 1])),
                                    sc2, nm2);
        const float2 e = make_float2(exp2f(x.x), exp2f(x.y));
        rs2 = __fadd2_rn(rs2, e);
        pv[i] = (kw & (1u << i)) ? e.x : 0.f;
        pv[i + 1] = (kw & (2u << i)) ? e.y : 0.f;
      }
      // P (bf16, dropped) goes straight back to TMEM as the A operand of the P V product: packed two keys per
      // column over S columns [0, 64), which every thread has read by now (the max exchange above is a block-wide
      // barrier after the S loads) — no shared-memory round trip for P
#pragma unroll
      for (int g8 = 0; g8 < 2; ++g8) {
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = pack_bf16(pv[16 * g8 + 2 * i], pv[16 * g8 + 2 * i + 1]);
        tmem_st8(tS + lane_sel + hf * 32 + kb * 16 + g8 * 8, pk);
      }
    }
    l_run = l_run * alpha + (rs2.x + rs2.y);
    m_run = m_next;
    // rescale the running output when some row of this warp raised its max (rare after tile 0)
    if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll 1
      for (int c = (hf == 0 ? 0 : 2); c < (hf == 0 ? 2 : 3); ++c) {
        uint32_t rr[32];
        tmem_ld32(tO + lane_sel + c * 32, rr);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) rr[i] = __float_as_uint(__uint_as_float(rr[i]) * alpha);
        tmem_st32(tO + lane_sel + c * 32, rr);
      }
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      mbar_wait(bar_v, j & 1);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 8; ++k)
        tc_mma_bf16_ts(tO, tS + 8 * k, step_mn64(bVm, k), kHi64, idesc_o, (j > 0 || k > 0) ? 1u : 0u);
      if (j + 1 < nkv) {  // the next S overwrites P: issued after (and executed in order behind) the P V MMAs
        mbar_wait(bar_k, (j + 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 6; ++k) tc_mma_bf16_lh(tS, step_k64(bQ, k), kHi64, step_k64(bK, k), kHi64, idesc_s, k > 0);
      }
      tc_commit(bar_mma);
    }
    __syncwarp();
  }
  if (nkv > 0) {
    mbar_wait(bar_mma, nkv & 1);
    tc_fence_after();
  }
  l_s[hf * 128 + r] = l_run;
  __syncthreads();
  {
    const float l_tot = l_s[r] + l_s[128 + r];
    const float inv_l = l_tot > 0.f ? p.inv_keep / l_tot : 0.f;
    const bool valid = q < p.Lq;
    // O / l -> bf16, staged in the (dead) Q tile and stored by one thread with TMA (rows past Lq are clipped)
#pragma unroll 1
    for (int c = (hf == 0 ? 0 : 2); c < (hf == 0 ? 2 : 3); ++c) {
      uint32_t rr[32];
      if (nkv > 0) {
        tmem_ld32(tO + lane_sel + c * 32, rr);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) rr[i] = 0u;
      }
      stage_chunk_sw64(sQ, r, c, rr, inv_l);
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      store_tile(&tmO, sQ, h * DH, q0, b);
      tma_commit_group();
      tma_wait_group_read0();
    }
    if (valid && hf == 0 && p.lse2)
      p.lse2[((long long)b * p.H + h) * p.Lq + q] = l_tot > 0.f ? (m_run + log2f(l_tot)) : INFINITY;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// =================================================================================================
// forward, second generation: the 128-key tile is processed as two 64-key halves so that, inside ONE CTA, the
// tensor pipe works on one half while the softmax warps work on the other (the first-generation kernel above
// alternates strictly between "MMA" and "softmax" per CTA and relies on the second resident CTA for overlap; its
// profile shows the softmax warps parked at barriers a third of the time).
//
//   warps 0-7   softmax: thread (r, hf) owns query row r and, in each half, keys [32 hf, 32 hf + 32)
//   warp 8      lane 0 issues every tcgen05.mma (all products accumulate into one O: they must come from one thread)
//   warp 9      lane 0 issues every TMA load (taken off the MMA thread: 20 MMAs + 12 bulk copies per tile at >= 54
//               cycles each made that one thread the critical path)
//   TMEM        S_a [0, 64)  S_b [64, 128)  O [128, 224);  P_x (packed bf16) overwrites the first 16 columns of
//               each thread's own 32 score columns, so no thread writes where another still has to read
//   smem        Q | K_a K_b (one 64-key half tile each) | V_a[2] V_b[2] (double-buffered halves)
//   per half x of tile j, in tensor-pipe order:  G(j, x) = [ O += P_x(j) V_x(j) ;  S_x(j + 1) = Q K_x(j + 1)^T ]
//   hand-offs   softmax -> issuer: bar_p[x] (8 warp arrivals: P_x(j) is in TMEM, S_x(j) has been read)
//               issuer -> softmax: bar_s[x] (tcgen05.commit after G(., x): S_x of the next tile is ready)
//   No block-wide barrier in the steady state.
//
// Running maximum without a per-tile exchange: the two threads of a row publish the maxima of the half tiles they have
// processed; the reference maximum m_ref used for tile j is decided at the start of tile j from what is guaranteed to be
// visible to both of them then (their half-a maxima through tile j - 1, half-b maxima through tile j - 2; the mbarrier
// chain through the issuer orders those writes), raised lazily (only when the seen maximum exceeds it by 2^8) and applied
// to both halves of the tile.  The probabilities of a tile can therefore exceed 1 — by the amount the scores grew within
// the last two tiles, harmless in fp32 / bf16 up to 2^100 — and the result O / l does not depend on the reference.
// When m_ref moves, each thread rescales its half of the O row after the last P V product issued so far has retired
// (it waits for that commit itself; the next product cannot be issued before this thread's own arrival).
// Tile 0 establishes the reference exactly (one block-wide exchange per CTA).
// =================================================================================================
constexpr int FWD2_THREADS = 320;
constexpr int HALF_TILE_BYTES = 64 * 192;  // [64 keys x 96] bf16 as three [64 x 32] swizzle-64 blocks of 4096 B
constexpr int FWD2_SMEM = 1024 + QKV_BYTES + 2 * HALF_TILE_BYTES + 4 * HALF_TILE_BYTES + 12 * 1024;

// one 64-row half tile of K or V by TMA (three 32-column blocks)
__device__ __forceinline__ void load_half(const CUtensorMap* m64, uint32_t bar, uint32_t dst, int col0, int row0, int b) {
  mbar_expect_tx(bar, HALF_TILE_BYTES);
#pragma unroll
  for (int c = 0; c < 3; ++c) tma_load_3d(m64, bar, dst + c * 4096, col0 + 32 * c, row0, b);
}

__global__ void __launch_bounds__(FWD2_THREADS, 2)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK64,
                 const __grid_constant__ CUtensorMap tmV64, const __grid_constant__ CUtensorMap tmO, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQ = base;
  const uint32_t sK = sQ + QKV_BYTES;                 // [2 halves]
  const uint32_t sV = sK + 2 * HALF_TILE_BYTES;       // [2 halves][2 stages]
  const uint32_t aux = sV + 4 * HALF_TILE_BYTES;
  float* mx_s = reinterpret_cast<float*>(gen + (aux - base));  // [2 halves][3 ring][2 hf][128] published half-tile maxima
  float* l_s = mx_s + 2 * 3 * 2 * 128;                          // [2][128]
  const uint32_t bars = aux + 2 * 3 * 2 * 128 * 4 + 2 * 128 * 4;
  const uint32_t bar_q = bars, bar_k = bars + 8 /*[2]*/, bar_v = bars + 24 /*[2 halves][2 stages]*/, bar_s = bars + 56 /*[2]*/,
                 bar_p = bars + 72 /*[2]*/;
  const uint32_t tmem_ptr_addr = bars + 96;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(gen + (tmem_ptr_addr - base));
  int* ext_slot = reinterpret_cast<int*>(gen + (bars + 104 - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nq_tiles = gridDim.x;
  const int qt = nq_tiles - 1 - blockIdx.x;  // heavy (late) causal tiles first
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * TILE;

  if (tid == 0) {
    mbar_init(bar_q, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_k + 8 * i, 1);
      mbar_init(bar_s + 8 * i, 1);
      mbar_init(bar_p + 8 * i, 8);
    }
    for (int i = 0; i < 4; ++i) mbar_init(bar_v + 8 * i, 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc(tmem_ptr_addr, 256);
  const int kv_end = kv_extent(p, b, ext_slot);  // contains __syncthreads when a mask is given
  int nkv = (kv_end + TILE - 1) / TILE;
  if (p.causal) nkv = min(nkv, qt + 1);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_gen;
  const uint32_t tO = tmem + 128;
  (void)lane;

  if (warp == 9) {
    // ===================== TMA producer (one thread) =====================
    if (lane == 0 && nkv > 0) {
      // prologue: Q, both K halves of tile 0, V halves of tiles 0 and 1; then K_x(1) as soon as S_x(0) has retired
      mbar_expect_tx(bar_q, QKV_BYTES);
      load_tile(&tmQ, bar_q, sQ, h * DH, q0, b);
      for (int x = 0; x < 2; ++x) load_half(&tmK64, bar_k + 8 * x, sK + x * HALF_TILE_BYTES, h * DH, 64 * x, b);
      for (int x = 0; x < 2; ++x) load_half(&tmV64, bar_v + 8 * (2 * x), sV + (2 * x) * HALF_TILE_BYTES, h * DH, 64 * x, b);
      if (nkv > 1) {
        for (int x = 0; x < 2; ++x)
          load_half(&tmV64, bar_v + 8 * (2 * x + 1), sV + (2 * x + 1) * HALF_TILE_BYTES, h * DH, TILE + 64 * x, b);
        for (int x = 0; x < 2; ++x) {
          mbar_wait(bar_s + 8 * x, 0);
          load_half(&tmK64, bar_k + 8 * x, sK + x * HALF_TILE_BYTES, h * DH, TILE + 64 * x, b);
        }
      }
      // buffers freed by G(j, x) (completion #(j + 1) of bar_s[x]): K_x (S_x(j + 1) retired) -> K_x(j + 2);
      // V_x stage j & 1 (the P V product of tile j retired) -> V_x(j + 2)
      for (int j = 0; j + 2 < nkv; ++j) {
        const int st = j & 1;
        for (int x = 0; x < 2; ++x) {
          mbar_wait(bar_s + 8 * x, (j + 1) & 1);
          load_half(&tmK64, bar_k + 8 * x, sK + x * HALF_TILE_BYTES, h * DH, (j + 2) * TILE + 64 * x, b);
          load_half(&tmV64, bar_v + 8 * (2 * x + st), sV + (2 * x + st) * HALF_TILE_BYTES, h * DH, (j + 2) * TILE + 64 * x, b);
        }
      }
    }
  } else if (warp == 8) {
    // ===================== MMA issuer (one thread: every product accumulates into the same O, and only the MMAs of
    //                       ONE thread are ordered among themselves) =====================
    if (lane == 0 && nkv > 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, false, false);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, DH, false, true);
      const uint32_t bQ = base_k64(sQ);
      auto issue_s = [&](int x) {  // S_x = Q K_x^T  (N = 64)
        const uint32_t bK = umma_desc_lo(sK + x * HALF_TILE_BYTES, 16);
#pragma unroll
        for (int k = 0; k < 6; ++k)
          tc_mma_bf16_lh(tmem + 64 * x, step_k64(bQ, k), kHi64, bK + ((((k >> 1) * 4096) + (k & 1) * 32) >> 4), kHi64,
                         idesc_s, k > 0);
      };
      auto issue_pv = [&](int x, int st, bool first) {  // O += P_x V_x  (K = 64 keys in four steps of 16)
        const uint32_t bV = umma_desc_lo(sV + (2 * x + st) * HALF_TILE_BYTES, 4096);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_bf16_ts(tO, tmem + 64 * x + 32 * (k >> 1) + 8 * (k & 1), bV + ((k * 1024) >> 4), kHi64, idesc_o,
                         (!first || k > 0) ? 1u : 0u);
      };
      mbar_wait(bar_q, 0);
      for (int x = 0; x < 2; ++x) {
        mbar_wait(bar_k + 8 * x, 0);
        tc_fence_after();
        issue_s(x);
        tc_commit(bar_s + 8 * x);  // completion #0 of bar_s[x]: S_x(0) ready
      }
      for (int j = 0; j < nkv; ++j) {
        const int st = j & 1;
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          mbar_wait(bar_p + 8 * x, j & 1);                       // P_x(j) written, S_x(j) consumed
          mbar_wait(bar_v + 8 * (2 * x + st), (j >> 1) & 1);     // V_x(j) landed
          tc_fence_after();
          issue_pv(x, st, j == 0 && x == 0);
          if (j + 1 < nkv) {
            mbar_wait(bar_k + 8 * x, (j + 1) & 1);               // K_x(j + 1) landed
            tc_fence_after();
            issue_s(x);
          }
          tc_commit(bar_s + 8 * x);                              // completion #(j + 1): G(j, x) retired
        }
      }
    }
  } else {
    // ===================== softmax warps =====================
    const int quad = warp & 3, hf = warp >> 2;
    const int r = quad * 32 + lane;  // local query row
    const int q = q0 + r;
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t bh = (uint32_t)(b * p.H + h);
    const DropKey dkey = drop_key(p);
    const uint8_t* mrow = p.kpm ? p.kpm + (long long)b * p.Lk : nullptr;
    float m_run = -INFINITY, m_seen = -INFINITY, l_run = 0.f;
    auto mx_at = [&](int x, int ring, int hh) -> float& { return mx_s[((x * 3 + ring) * 2 + hh) * 128 + r]; };
    // bit i set <=> key key0 + i takes no part (padding mask, past Lk, above the causal diagonal).  Zero for almost every
    // block: the 32 mask bytes are OR-ed as two 16-byte words first.
    auto dead_bits = [&](int j, int x) -> uint32_t {
      const int kloc = 64 * x + 32 * hf;  // first key of the block inside the tile
      const int key0 = j * TILE + kloc;
      uint32_t dead = 0u;
      if (key0 + 32 > p.Lk) dead = key0 >= p.Lk ? 0xFFFFFFFFu : (0xFFFFFFFFu << (p.Lk - key0));
      if (mrow != nullptr && key0 < p.Lk) {
        const int n = min(32, p.Lk - key0);
        if (n == 32 && (reinterpret_cast<uintptr_t>(mrow + key0) & 15) == 0) {
          const uint4 a = *reinterpret_cast<const uint4*>(mrow + key0);
          const uint4 c = *reinterpret_cast<const uint4*>(mrow + key0 + 16);
          if ((a.x | a.y | a.z | a.w | c.x | c.y | c.z | c.w) != 0u) {
            const uint32_t w[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if ((w[i >> 2] >> (8 * (i & 3))) & 0xFFu) dead |= 1u << i;
          }
        } else {
          for (int i = 0; i < n; ++i)
            if (mrow[key0 + i] != 0) dead |= 1u << i;
        }
      }
      if (p.causal && j == qt) {  // keys kloc + i > r
        const int first_dead = r + 1 - kloc;  // index of the first key above the diagonal
        if (first_dead <= 0) dead = 0xFFFFFFFFu;
        else if (first_dead < 32) dead |= 0xFFFFFFFFu << first_dead;
      }
      return dead;
    };

    // one half tile: 32 scores of this thread -> P (packed bf16, back into the first 16 of its own score columns);
    // returns the maximum of the (masked) scaled scores
    auto half_tile = [&](int j, int x, float m_eff) -> float {
      const int key0 = j * TILE + 64 * x + 32 * hf;  // first key of this thread's 32
      const uint32_t t_s = tmem + lane_sel + 64 * x + 32 * hf;
      const uint32_t dead = dead_bits(j, x);  // index-only work: before the scores are waited for
      const uint32_t kw = keep_word(p, dkey, bh, (uint32_t)q, (uint32_t)(key0 >> 5));
      mbar_wait(bar_s + 8 * x, j & 1);
      tc_fence_after();
      uint32_t sr[32];
      tmem_ld32(t_s, sr);
      tmem_ld_wait();
      if (dead != 0u) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (dead & (1u << i)) sr[i] = __float_as_uint(-INFINITY);
      }
      float mx = __uint_as_float(sr[0]);
#pragma unroll
      for (int i = 1; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(sr[i]));
      const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(-m_eff, -m_eff);
      float2 rs2 = make_float2(0.f, 0.f);
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float2 xx = __ffma2_rn(make_float2(__uint_as_float(sr[i]), __uint_as_float(sr[i + 1])), sc2, nm2);
        const float2 e = make_float2(exp2f(xx.x), exp2f(xx.y));
        rs2 = __fadd2_rn(rs2, e);
        pk[i >> 1] = pack_bf16((kw & (1u << i)) ? e.x : 0.f, (kw & (2u << i)) ? e.y : 0.f);
      }
      l_run += rs2.x + rs2.y;
      {
        const uint32_t(&p0)[8] = *reinterpret_cast<const uint32_t(*)[8]>(&pk[0]);
        const uint32_t(&p1)[8] = *reinterpret_cast<const uint32_t(*)[8]>(&pk[8]);
        tmem_st8(t_s, p0);
        tmem_st8(t_s + 8, p1);
      }
      return mx * p.scale_log2;  // scale > 0
    };
    auto arrive_p = [&](int x) {
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p + 8 * x);
    };

    for (int j = 0; j < nkv; ++j) {
      const int ring = j % 3;
      // ---- reference maximum of this tile
      float m_next = m_run;
      if (j == 0) {
        // exact maximum of tile 0: both halves' scores, both threads of the row (one block-wide exchange per CTA)
        float mx = -INFINITY;
#pragma unroll 1
        for (int x = 0; x < 2; ++x) {
          mbar_wait(bar_s + 8 * x, 0);
          tc_fence_after();
          uint32_t sr[32];
          tmem_ld32(tmem + lane_sel + 64 * x + 32 * hf, sr);
          tmem_ld_wait();
          const uint32_t dead = dead_bits(0, x);
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (!(dead & (1u << i))) mx = fmaxf(mx, __uint_as_float(sr[i]));
        }
        l_s[hf * 128 + r] = mx;
        named_bar_sync(1, 256);
        m_seen = fmaxf(mx, l_s[(hf ^ 1) * 128 + r]) * p.scale_log2;
        named_bar_sync(1, 256);  // l_s is reused at the end
        m_next = m_seen;
      } else {
        // what both threads of the row are guaranteed to see by now: half a of tile j - 1, half b of tile j - 2
        const int ra = (j - 1) % 3;
        m_seen = fmaxf(m_seen, fmaxf(mx_at(0, ra, 0), mx_at(0, ra, 1)));
        if (j >= 2) {
          const int rb = (j - 2) % 3;
          m_seen = fmaxf(m_seen, fmaxf(mx_at(1, rb, 0), mx_at(1, rb, 1)));
        }
        if (m_run == -INFINITY || m_seen > m_run + 8.0f) m_next = fmaxf(m_run, m_seen);
      }
      {
        // rescale l and this thread's half of the O row (rare after the first tiles).  Warp-uniform: tcgen05.ld / st are
        // .sync.aligned, so every lane takes part as soon as one row of the warp moves its reference (alpha = 1 for
        // the others).
        const bool moved = m_next != m_run && m_run != -INFINITY;
        if (__any_sync(0xffffffffu, moved)) {
          const float alpha = moved ? exp2f(m_run - m_next) : 1.0f;
          l_run *= alpha;
          mbar_wait(bar_s + 8, j & 1);  // G(j - 1, b) retired: no P V product is in flight or can be issued
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 3; ++c) {
            uint32_t rr[16];
            tmem_ld16(tO + lane_sel + 48 * hf + 16 * c, rr);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) rr[i] = __float_as_uint(__uint_as_float(rr[i]) * alpha);
            tmem_st16(tO + lane_sel + 48 * hf + 16 * c, rr);
          }
          tmem_st_wait();
        }
        m_run = m_next;
      }
      const float m_eff = (m_run == -INFINITY) ? 0.f : m_run;
#pragma unroll 1
      for (int x = 0; x < 2; ++x) {
        const float mx = half_tile(j, x, m_eff);
        mx_at(x, ring, hf) = mx;  // published before the arrival below (the issuer's acquire orders it)
        arrive_p(x);
      }
    }
    // ---- epilogue
    if (nkv > 0) {
      mbar_wait(bar_s, nkv & 1);      // G(nkv - 1, a)
      mbar_wait(bar_s + 8, nkv & 1);  // G(nkv - 1, b)
      tc_fence_after();
    }
    l_s[hf * 128 + r] = l_run;
    named_bar_sync(1, 256);
    const float l_tot = l_s[r] + l_s[128 + r];
    const float inv_l = l_tot > 0.f ? p.inv_keep / l_tot : 0.f;
    // O / l -> bf16, staged in the (dead) Q tile and stored by one thread with TMA (rows past Lq are clipped)
#pragma unroll 1
    for (int c = (hf == 0 ? 0 : 2); c < (hf == 0 ? 2 : 3); ++c) {
      uint32_t rr[32];
      if (nkv > 0) {
        tmem_ld32(tO + lane_sel + c * 32, rr);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) rr[i] = 0u;
      }
      stage_chunk_sw64(sQ, r, c, rr, inv_l);
    }
    fence_proxy_async_smem();
    named_bar_sync(1, 256);
    if (tid == 0) {
      store_tile(&tmO, sQ, h * DH, q0, b);
      tma_commit_group();
      tma_wait_group_read0();
    }
    if (q < p.Lq && hf == 0 && p.lse2)
      p.lse2[((long long)b * p.H + h) * p.Lq + q] = l_tot > 0.f ? (m_run + log2f(l_tot)) : INFINITY;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 256);
}

// =================================================================================================
// decode step (Lq = 1, inference): one block per (head, batch) streams that head's K and V rows once — the op is a
// pure read of the KV cache (2 * Lk * 192 B per block), so it is written as a SIMT stream, not as 128-row tensor-core
// tiles of which one row would be real.  Four lanes share a key row (48 bytes = 24 head dims = three 16-byte loads
// each, so a warp reads eight whole 192-byte rows per load instruction: coalesced); each 4-lane group keeps a private
// online softmax (running max, sum, 24 output accumulators per lane) over keys g, g + 32, ..., and the groups of the
// block are rescaled to the common maximum and folded through shared memory at the end.
// =================================================================================================
template <int DEC_THREADS>
__global__ void __launch_bounds__(DEC_THREADS)
attn_decode_kernel(const __nv_bfloat16* __restrict__ q, long long ldq, const __nv_bfloat16* __restrict__ k,
                   const __nv_bfloat16* __restrict__ v, long long ldkv, long long kv_bs, __nv_bfloat16* __restrict__ o,
                   long long ldo, float* __restrict__ lse2, const uint8_t* __restrict__ kpm, int H, int Lk,
                   float scale_log2) {
  constexpr int LPK = 4;             // lanes per key row: 48 bytes = 24 head dims = three 16-byte loads each
  constexpr int DPL = DH / LPK;      // head dims per lane
  constexpr int DEC_GROUPS = DEC_THREADS / LPK;
  __shared__ float part[DEC_GROUPS][DH];
  __shared__ float ms[DEC_GROUPS], ls[DEC_GROUPS];
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int g = tid / LPK, j = tid % LPK;  // key group, slice of the head
  const unsigned gmask = 0xFu << (tid & 28);
  float qr[DPL];
  {
    const uint4* qp = reinterpret_cast<const uint4*>(q + (long long)b * ldq + h * DH + j * DPL);
#pragma unroll
    for (int c = 0; c < DPL / 8; ++c) {
      const uint4 u = qp[c];
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = unpack_bf16(w[i]);
        qr[8 * c + 2 * i] = f.x * scale_log2;  // scores come out in the log2 domain
        qr[8 * c + 2 * i + 1] = f.y * scale_log2;
      }
    }
  }
  const __nv_bfloat16* kb = k + (long long)b * kv_bs + h * DH + j * DPL;
  const __nv_bfloat16* vb = v + (long long)b * kv_bs + h * DH + j * DPL;
  const uint8_t* mrow = kpm ? kpm + (long long)b * Lk : nullptr;
  float m = -INFINITY, l = 0.f;
  float acc[DPL];
#pragma unroll
  for (int i = 0; i < DPL; ++i) acc[i] = 0.f;
#pragma unroll 2
  for (int kk = g; kk < Lk; kk += DEC_GROUPS) {
    if (mrow != nullptr && mrow[kk] != 0) continue;  // uniform inside the group
    const uint4* krow = reinterpret_cast<const uint4*>(kb + (long long)kk * ldkv);
    const uint4* vrow = reinterpret_cast<const uint4*>(vb + (long long)kk * ldkv);
    uint4 kr[DPL / 8], vr[DPL / 8];
#pragma unroll
    for (int c = 0; c < DPL / 8; ++c) kr[c] = krow[c];
#pragma unroll
    for (int c = 0; c < DPL / 8; ++c) vr[c] = vrow[c];
    float sc = 0.f;
#pragma unroll
    for (int c = 0; c < DPL / 8; ++c) {
      const uint32_t w[4] = {kr[c].x, kr[c].y, kr[c].z, kr[c].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = unpack_bf16(w[i]);
        sc += f.x * qr[8 * c + 2 * i] + f.y * qr[8 * c + 2 * i + 1];
      }
    }
    sc += __shfl_xor_sync(gmask, sc, 1);  // the groups of a warp run different trip counts: group-wide mask
    sc += __shfl_xor_sync(gmask, sc, 2);
    const float m_new = fmaxf(m, sc);
    const float alpha = exp2f(m - m_new);  // first key: exp2(-inf) = 0
    const float e = exp2f(sc - m_new);
    l = l * alpha + e;
    m = m_new;
#pragma unroll
    for (int c = 0; c < DPL / 8; ++c) {
      const uint32_t w[4] = {vr[c].x, vr[c].y, vr[c].z, vr[c].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = unpack_bf16(w[i]);
        acc[8 * c + 2 * i] = acc[8 * c + 2 * i] * alpha + e * f.x;
        acc[8 * c + 2 * i + 1] = acc[8 * c + 2 * i + 1] * alpha + e * f.y;
      }
    }
  }
  if (j == 0) ms[g] = m;
  __syncthreads();
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < DEC_GROUPS; ++i) mx = fmaxf(mx, ms[i]);
  const float rs = (m == -INFINITY) ? 0.f : exp2f(m - mx);  // this group's weight (0 if it saw no key)
  if (j == 0) ls[g] = l * rs;
#pragma unroll
  for (int i = 0; i < DPL; ++i) part[g][j * DPL + i] = acc[i] * rs;
  __syncthreads();
  if (tid < DH) {
    float lt = 0.f, st = 0.f;
#pragma unroll
    for (int i = 0; i < DEC_GROUPS; ++i) {
      lt += ls[i];
      st += part[i][tid];
    }
    o[(long long)b * ldo + h * DH + tid] = __float2bfloat16(lt > 0.f ? st / lt : 0.f);
    if (tid == 0 && lse2 != nullptr) lse2[(long long)b * H + h] = lt > 0.f ? mx + log2f(lt) : INFINITY;
  }
}

// =================================================================================================
// backward preprocess: D[b,h,q] = sum_c dO[b,q,h,c] * O[b,q,h,c]
// =================================================================================================
__global__ void __launch_bounds__(256)
attn_bwd_dvec_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                     float* __restrict__ dvec, int rows, int Lq, int H, int ldo) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int b = row / Lq, q = row % Lq;
  // lane covers 24 consecutive elements of the 768-wide row; 4 lanes = one head (dh = 96)
  const __nv_bfloat16* po = o + (long long)row * ldo + lane * 24;
  const __nv_bfloat16* pg = d_o + (long long)row * ldo + lane * 24;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const uint4 a = *reinterpret_cast<const uint4*>(po + i * 8);
    const uint4 g = *reinterpret_cast<const uint4*>(pg + i * 8);
    const uint32_t aa[4] = {a.x, a.y, a.z, a.w}, gg[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 x = unpack_bf16(aa[k]), y = unpack_bf16(gg[k]);
      s += x.x * y.x + x.y * y.y;
    }
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  if ((lane & 3) == 0) {
    const int h = lane >> 2;
    if (h < H) dvec[((long long)b * H + h) * Lq + q] = s;
  }
}

// =================================================================================================
// Backward kernels: warp-specialised and software-pipelined.  A TMA producer warp, a single-lane tcgen05
// issuer warp and two arithmetic warpgroups that only meet through mbarriers (no block-wide barrier in
// the loops).  The 128x128 score tile is processed as two 128x64 halves, each owned by one warpgroup with
// its own TMEM buffers, so the tensor pipe computes one half (and the gradient GEMMs of the previous one)
// while the other half is in the exp / dropout / dS arithmetic.
//   dQ kernel   : one CTA per (q-tile, head, batch), rows = queries, key tiles stream through
//   dK/dV kernel: one CTA per (key tile, head, batch), rows = keys, query tiles stream through; the dropout
//                 keep-bits (generated per query row) are transposed 32x32 inside the warp
// Measured on B200 (tools/attn_bench.py, B=32 H=8 L=1024): 8 arithmetic warps beat 16 (launch-/latency-
// bound hand-offs, not issue-bound), and this version beats the earlier phase-serial kernels by ~7 %.
// =================================================================================================
constexpr int HALF_BYTES = TILE * 128;   // one [128 x 64] bf16 swizzle-128 block (P^T / dS halves)
constexpr int BWD3_THREADS = 320;  // 8 arithmetic warps + MMA warp + TMA warp

// K-major [128 x 64] bf16 swizzle-128 block, k16 step k (0..3)
__device__ __forceinline__ uint32_t lo_k128h(uint32_t tile, int k) { return umma_desc_lo(tile + k * 32, 16); }

// ---- dQ: one CTA per (q-tile, head, batch); key tiles stream through, each split in two 64-key halves ----
//   S_h = Q K_h^T, dP_h = dO V_h^T (TMEM, rows = q) -> dS_h (bf16, smem) -> dQ += dS_h K_h
constexpr int BWD3_Q_SMEM = 1024 + 2 * QKV_BYTES + 4 * QKV_BYTES + 2 * HALF_BYTES + 1024;

__global__ void __launch_bounds__(BWD3_THREADS, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                    const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQ = base, sdO = sQ + QKV_BYTES;
  const uint32_t sK = sdO + QKV_BYTES;           // [2] stages
  const uint32_t sV = sK + 2 * QKV_BYTES;        // [2] stages
  const uint32_t sDS = sV + 2 * QKV_BYTES;       // [2] halves
  const uint32_t aux = sDS + 2 * HALF_BYTES;
  uint8_t* mask_s = gen + (aux - base);          // [2][128]
  int* anym_s = reinterpret_cast<int*>(gen + (aux + 256 - base));
  int* ext_slot = reinterpret_cast<int*>(gen + (aux + 272 - base));
  const uint32_t bar0 = aux + 512;
  const uint32_t bar_q = bar0, bar_kvf = bar0 + 8, bar_kve = bar0 + 24, bar_mkf = bar0 + 40, bar_sf = bar0 + 56,
                 bar_pf = bar0 + 72, bar_dsd = bar0 + 88, bar_fin = bar0 + 104;
  const uint32_t tmem_ptr_addr = bar0 + 112;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(gen + (tmem_ptr_addr - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nq_tiles = gridDim.x;
  const int qt = nq_tiles - 1 - blockIdx.x;
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * TILE;

  if (tid == 0) {
    mbar_init(bar_q, 1);
    mbar_init(bar_fin, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_kvf + 8 * i, 1);
      mbar_init(bar_kve + 8 * i, 1);
      mbar_init(bar_mkf + 8 * i, 1);
      mbar_init(bar_sf + 8 * i, 1);
      mbar_init(bar_pf + 8 * i, 4);
      mbar_init(bar_dsd + 8 * i, 1);
    }
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc(tmem_ptr_addr, 512);
  int nkv = (kv_extent(p, b, ext_slot) + TILE - 1) / TILE;
  if (p.causal) nkv = min(nkv, qt + 1);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_gen;
  // TMEM columns: S_h at 64h, dP_h at 128 + 64h, dQ at 256
  if (warp == 9) {
    if (lane == 0 && nkv > 0) {
      mbar_expect_tx(bar_q, 2 * QKV_BYTES);
      load_tile(&tmQ, bar_q, sQ, h * DH, q0, b);
      load_tile(&tmdO, bar_q, sdO, h * DH, q0, b);
    }
    for (int j = 0; j < nkv; ++j) {
      const int st = j & 1, u = j >> 1;
      if (u > 0) mbar_wait(bar_kve + 8 * st, (u - 1) & 1);
      if (lane == 0) {
        mbar_expect_tx(bar_kvf + 8 * st, 2 * QKV_BYTES);
        load_tile(&tmK, bar_kvf + 8 * st, sK + st * QKV_BYTES, h * DH, j * TILE, b);
        load_tile(&tmV, bar_kvf + 8 * st, sV + st * QKV_BYTES, h * DH, j * TILE, b);
      }
      uint32_t any = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = lane * 4 + i, kv = j * TILE + c;
        uint8_t m = kv >= p.Lk ? 1 : 0;
        if (!m && p.kpm) m = p.kpm[(long long)b * p.Lk + kv] != 0;
        mask_s[st * 128 + c] = m;
        any |= m;
      }
      any = __any_sync(0xffffffffu, any != 0);
      if (lane == 0) anym_s[st] = any;
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_mkf + 8 * st);
    }
  } else if (warp == 8) {
    if (lane == 0 && nkv > 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, false, false);
      constexpr uint32_t idesc_g = umma_idesc_bf16(128, DH, false, true);
      auto issue_s = [&](int hh, int st) {  // S_h and dP_h of the key tile staged in `st`
        const uint32_t kh = sK + st * QKV_BYTES + hh * 4096, vh = sV + st * QKV_BYTES + hh * 4096;
#pragma unroll
        for (int k = 0; k < 6; ++k) tc_mma_bf16_lh(tmem + 64 * hh, lo_k64(sQ, k), kHi64, lo_k64(kh, k), kHi64, idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < 6; ++k) tc_mma_bf16_lh(tmem + 128 + 64 * hh, lo_k64(sdO, k), kHi64, lo_k64(vh, k), kHi64, idesc_s, k > 0);
        tc_commit(bar_sf + 8 * hh);
      };
      mbar_wait(bar_q, 0);
      mbar_wait(bar_kvf, 0);
      tc_fence_after();
      issue_s(0, 0);
      issue_s(1, 0);
      for (int j = 0; j < nkv; ++j) {
        const int st = j & 1;
        if (j + 1 < nkv) {
          mbar_wait(bar_kvf + 8 * (st ^ 1), ((j + 1) >> 1) & 1);
          tc_fence_after();
        }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          mbar_wait(bar_pf + 8 * hh, j & 1);  // dS_h(j) in smem; S_h / dP_h consumed
          tc_fence_after();
          const uint32_t kh = sK + st * QKV_BYTES + hh * 4096;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_bf16_lh(tmem + 256, lo_k128h(sDS + hh * HALF_BYTES, k), kHi128, lo_mn64(kh, k), kHi64, idesc_g,
                        (j > 0 || hh > 0 || k > 0) ? 1u : 0u);
          tc_commit(bar_dsd + 8 * hh);
          if (j + 1 < nkv) issue_s(hh, st ^ 1);
        }
        tc_commit(bar_kve + 8 * st);
      }
      tc_commit(bar_fin);  // every dQ MMA has retired (a parity wait on the other half's barrier could pass early)
    }
  } else {
    const int hh = warp >> 2, quad = warp & 3;  // warpgroup hh owns key columns [64 hh, 64 hh + 64) of every tile
    const int r = quad * 32 + lane;
    const int q = q0 + r;
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem + 64 * hh, tDP = tmem + 128 + 64 * hh;
    const uint32_t sDSh = sDS + hh * HALF_BYTES;
    const long long stat_o = ((long long)b * p.H + h) * p.Lq + q;
    const float lse2 = q < p.Lq ? p.lse2[stat_o] : INFINITY;
    const float dvec = q < p.Lq ? p.dvec[stat_o] : 0.f;
    const uint32_t bh = (uint32_t)(b * p.H + h);
    const DropKey dkey = drop_key(p);
    for (int j = 0; j < nkv; ++j) {
      const int st = j & 1;
      const int kv0 = j * TILE + hh * 64;
      const uint32_t kw0 = keep_word(p, dkey, bh, (uint32_t)q, (uint32_t)(kv0 >> 5));  // index-only: before the waits
      const uint32_t kw1 = keep_word(p, dkey, bh, (uint32_t)q, (uint32_t)(kv0 >> 5) + 1);
      mbar_wait(bar_mkf + 8 * st, (j >> 1) & 1);
      const bool any_mask = anym_s[st] != 0;
      mbar_wait(bar_sf + 8 * hh, j & 1);
      tc_fence_after();
      const bool diag = p.causal && (j == qt);
      if (j > 0) mbar_wait(bar_dsd + 8 * hh, (j - 1) & 1);  // dS_h buffer free again
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c0 = cc * 32;
        const uint32_t kw = cc == 0 ? kw0 : kw1;
        uint32_t rs[32], rp[32];
        tmem_ld32(tS + lane_sel + c0, rs);
        tmem_ld32(tDP + lane_sel + c0, rp);
        tmem_ld_wait();
        float ds[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int c = hh * 64 + c0 + i;  // key column inside the 128-key tile
          float pr = exp2f(fmaf(__uint_as_float(rs[i]), p.scale_log2, -lse2));
          if (any_mask) pr = mask_s[st * 128 + c] ? 0.f : pr;
          if (diag && c > r) pr = 0.f;
          const float dm = (kw & (1u << i)) ? p.inv_keep : 0.f;
          ds[i] = pr * fmaf(__uint_as_float(rp[i]), dm, -dvec);
        }
        store_row32_sw128(sDSh, r, c0, ds);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_pf + 8 * hh);
    }
    // epilogue: dQ * scale -> bf16; warpgroup 0 stores columns [0,64), warpgroup 1 columns [64,96)
    if (nkv > 0) {
      mbar_wait(bar_fin, 0);  // the last gradient GEMM
      tc_fence_after();
    }
    {
      __nv_bfloat16* dst = p.dq + ((long long)b * p.Lq + q) * p.ldq_out + h * DH;
#pragma unroll 1
      for (int c = hh * 2; c < (hh == 0 ? 2 : 3); ++c) {
        uint32_t rr[32];
        if (nkv > 0) {
          tmem_ld32(tmem + 256 + lane_sel + c * 32, rr);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) rr[i] = 0u;
        }
        if (q < p.Lq) {
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) {
            uint4 u;
            u.x = pack_bf16(__uint_as_float(rr[8 * qd]) * p.scale, __uint_as_float(rr[8 * qd + 1]) * p.scale);
            u.y = pack_bf16(__uint_as_float(rr[8 * qd + 2]) * p.scale, __uint_as_float(rr[8 * qd + 3]) * p.scale);
            u.z = pack_bf16(__uint_as_float(rr[8 * qd + 4]) * p.scale, __uint_as_float(rr[8 * qd + 5]) * p.scale);
            u.w = pack_bf16(__uint_as_float(rr[8 * qd + 6]) * p.scale, __uint_as_float(rr[8 * qd + 7]) * p.scale);
            *reinterpret_cast<uint4*>(dst + c * 32 + qd * 8) = u;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 512);
}

// One [128 keys x 64 queries] half tile of the dK/dV kernel's arithmetic: S^T and dP^T (TMEM, fp32) -> P^T (packed bf16,
// back to TMEM: A operand of dV += P^T dO) and dS^T (bf16, shared memory: A operand of dK += dS^T Q and source of the
// workspace store).  Thread = key row r; four 16-query chunks, the TMEM loads of chunk cc + 1 in flight while chunk cc
// is in the ALUs.  Per score: P = exp2(s * c - lse), m = keep ? 1 / (1 - p) : 0, P~ = P * m, dS = P * (dP * m - D) as
// packed fp32 pairs (FFMA2 / FMUL2: ~5 instructions per score instead of ~13 scalar ones; the statistics arrive
// NEGATED in shared memory so they are plain FFMA2 addends).
// NSUB = 1: one thread per key row does all four chunks (8 arithmetic warps); NSUB = 2: two threads per key row, in
// warps of the same TMEM lane quadrant, take two chunks each (`sub` = 0 / 1; 16 arithmetic warps: four per scheduler
// instead of two, half the live registers per thread; kw0 is then the keep word of this thread's 32 queries).
template <bool DIAG, bool ROWMASK, int NSUB>
__device__ __forceinline__ void dkdv_half_tile(uint32_t tST, uint32_t tDPT, uint32_t tPT, uint32_t lane_sel,
                                               uint32_t sDSTh, const float* __restrict__ nlse,
                                               const float* __restrict__ ndv, uint32_t kw0, uint32_t kw1, int r, int hh,
                                               bool row_valid, float scale_log2, float inv_keep, int sub,
                                               uint32_t bar_consumed) {
  constexpr int NCH = 4 / NSUB;  // chunks per thread
  const int cc0 = NSUB == 1 ? 0 : 2 * sub;
  uint32_t rs[2][16], rp[2][16];
  tmem_ld16(tST + lane_sel + cc0 * 16, rs[0]);
  tmem_ld16(tDPT + lane_sel + cc0 * 16, rp[0]);
  if (NSUB == 2) {  // both chunks of this thread at once: its share of S^T / dP^T is in registers after ONE wait
    tmem_ld16(tST + lane_sel + cc0 * 16 + 16, rs[1]);
    tmem_ld16(tDPT + lane_sel + cc0 * 16 + 16, rp[1]);
  }
  // `bar_consumed`: S^T_h / dP^T_h have been read — the issuer may overwrite them with the NEXT query tile's scores
  // while this tile is still in the ALUs (the profile showed the arithmetic warps waiting for those MMAs 31 % of
  // the time when the hand-off only happened after the whole tile had been processed)
  auto signal_consumed = [&]() {
    tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar_consumed);
  };
  const float2 sc2 = make_float2(scale_log2, scale_log2);
#pragma unroll
  for (int ci = 0; ci < NCH; ++ci) {
    const int cc = cc0 + ci;
    const int c0 = cc * 16;
    const uint32_t kw = NSUB == 1 ? ((cc < 2 ? kw0 : kw1) >> ((cc & 1) * 16)) : (kw0 >> (ci * 16));
    if (NSUB == 2) {
      if (ci == 0) {
        tmem_ld_wait();
        signal_consumed();
      }
    } else {
      tmem_ld_wait();
      if (ci + 1 < NCH) {
        tmem_ld16(tST + lane_sel + c0 + 16, rs[(ci + 1) & 1]);
        tmem_ld16(tDPT + lane_sel + c0 + 16, rp[(ci + 1) & 1]);
      } else {
        signal_consumed();
      }
    }
    const uint32_t (&s_)[16] = rs[ci & 1];
    const uint32_t (&p_)[16] = rp[ci & 1];
    float pd[16], ds[16];
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      const float2 nl = *reinterpret_cast<const float2*>(nlse + c0 + i);
      const float2 nd = *reinterpret_cast<const float2*>(ndv + c0 + i);
      const float2 x = __ffma2_rn(make_float2(__uint_as_float(s_[i]), __uint_as_float(s_[i + 1])), sc2, nl);
      float2 pr = make_float2(exp2f(x.x), exp2f(x.y));
      if (DIAG) {  // causal: key r sees queries >= r only
        if (r > hh * 64 + c0 + i) pr.x = 0.f;
        if (r > hh * 64 + c0 + i + 1) pr.y = 0.f;
      }
      if (ROWMASK) {
        pr.x = row_valid ? pr.x : 0.f;
        pr.y = row_valid ? pr.y : 0.f;
      }
      const float2 dm = make_float2((kw & (1u << i)) ? inv_keep : 0.f, (kw & (2u << i)) ? inv_keep : 0.f);
      const float2 u = __ffma2_rn(make_float2(__uint_as_float(p_[i]), __uint_as_float(p_[i + 1])), dm, nd);
      const float2 d = __fmul2_rn(pr, u);
      const float2 pk = __fmul2_rn(pr, dm);
      ds[i] = d.x;
      ds[i + 1] = d.y;
      pd[i] = pk.x;
      pd[i + 1] = pk.y;
    }
    {  // P^T chunk -> TMEM (8 packed columns per 16 queries): the A operand of the dV MMA, no shared-memory trip
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) pk[i] = pack_bf16(pd[2 * i], pd[2 * i + 1]);
      tmem_st8(tPT + lane_sel + cc * 8, pk);
    }
    store_row16_sw128(sDSTh, r, c0, ds);  // dS^T chunk -> shared memory: A operand of the dK MMA and workspace store
  }
}

// ---- dK/dV: one CTA per (key tile, head, batch); query tiles stream through in 64-query halves ----
//   S^T_h = K Q_h^T, dP^T_h = V dO_h^T (TMEM, rows = keys) -> P^T_h, dS^T_h (bf16, smem)
//   -> dV += P^T_h dO_h, dK += dS^T_h Q_h
constexpr int BWD3_KV_SMEM = 1024 + 2 * QKV_BYTES + 4 * QKV_BYTES + 4 * HALF_BYTES + 4096;

// Warps: 0-7 arithmetic (two warpgroups, one per 64-query half), 8 = S^T / dP^T MMA issuer, 9 = TMA producer,
// 10 = dV / dK MMA issuer.  TWO issuing threads on different schedulers: a single thread needs >= 54 clk per
// tcgen05.mma (tools/micro/mma_issue.cu) and ~70-90 clk when it shares its scheduler with busy arithmetic warps, which
// made the one issuer of 40 small MMAs per query tile the critical path (measured with clock64 traces).
// AW = 16 (two threads per key row): 19 warps, <= 104 registers per thread.
constexpr int bwd_kv_threads(int aw) { return (aw + 3) * 32; }
template <int AW>
__global__ void __launch_bounds__(bwd_kv_threads(AW), 1)
attn_bwd_dkdv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                      const __grid_constant__ CUtensorMap tmDS, const __grid_constant__ CUtensorMap tmDK,
                      const __grid_constant__ CUtensorMap tmDV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sK = base, sV = sK + QKV_BYTES;
  const uint32_t sQ = sV + QKV_BYTES;            // [2] stages
  const uint32_t sdO = sQ + 2 * QKV_BYTES;       // [2] stages
  const uint32_t sPT = sdO + 2 * QKV_BYTES;      // [2] halves
  const uint32_t sDST = sPT + 2 * HALF_BYTES;    // [2] halves
  const uint32_t aux = sDST + 2 * HALF_BYTES;
  float* lse_s = reinterpret_cast<float*>(gen + (aux - base));  // [2][128]
  float* dv_s = lse_s + 256;                                     // [2][128]
  int* ext_slot = reinterpret_cast<int*>(gen + (aux + 2048 - base));
  const uint32_t bar0 = aux + 2560;
  const uint32_t bar_kv = bar0, bar_qf = bar0 + 8, bar_qe = bar0 + 24, bar_stf = bar0 + 40, bar_sf = bar0 + 56,
                 bar_pf = bar0 + 72, bar_gd = bar0 + 88, bar_fin = bar0 + 104;
  const uint32_t tmem_ptr_addr = bar0 + 112;
  const uint32_t bar_sc = bar0 + 128;  // [2] S^T_h / dP^T_h of the current tile are in the arithmetic warps' registers
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(gen + (tmem_ptr_addr - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int jt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int kv0 = jt * TILE;
  const int nq_tiles = (p.Lq + TILE - 1) / TILE;
  const int i_begin = p.causal ? jt : 0;

  if (tid == 0) {
    mbar_init(bar_kv, 1);
    mbar_init(bar_fin, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_qf + 8 * i, 1);
      mbar_init(bar_qe + 8 * i, 2);  // both MMA issuers release a Q / dO stage
      mbar_init(bar_stf + 8 * i, 1);
      mbar_init(bar_sf + 8 * i, 1);
      mbar_init(bar_pf + 8 * i, AW / 2);  // one arrival per arithmetic warp of the half
      mbar_init(bar_sc + 8 * i, AW / 2);
      mbar_init(bar_gd + 8 * i, 1);
    }
    fence_mbar_init();
  }
  constexpr int W_S = AW, W_TMA = AW + 1, W_G = AW + 2;  // S^T / dP^T issuer, TMA producer, dV / dK issuer
  if (warp == W_S) tmem_alloc(tmem_ptr_addr, 512);
  const int n_it = (kv0 < kv_extent(p, b, ext_slot)) ? nq_tiles - i_begin : 0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_gen;
  // TMEM columns: S^T_h at 64h, dP^T_h at 128 + 64h, dV at 256, dK at 352, P^T_h (packed bf16) at 448 + 32h
  if (warp == W_TMA) {
    if (lane == 0 && n_it > 0) {
      mbar_expect_tx(bar_kv, 2 * QKV_BYTES);
      load_tile(&tmK, bar_kv, sK, h * DH, kv0, b);
      load_tile(&tmV, bar_kv, sV, h * DH, kv0, b);
    }
    for (int it = 0; it < n_it; ++it) {
      const int st = it & 1, u = it >> 1;
      const int q0 = (i_begin + it) * TILE;
      if (u > 0) mbar_wait(bar_qe + 8 * st, (u - 1) & 1);
      if (lane == 0) {
        mbar_expect_tx(bar_qf + 8 * st, 2 * QKV_BYTES);
        load_tile(&tmQ, bar_qf + 8 * st, sQ + st * QKV_BYTES, h * DH, q0, b);
        load_tile(&tmdO, bar_qf + 8 * st, sdO + st * QKV_BYTES, h * DH, q0, b);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = lane * 4 + i, qq = q0 + c;
        const long long o = ((long long)b * p.H + h) * p.Lq + qq;
        // negated: the arithmetic warps use them as the addend of packed FFMA2s (exp2(s * c - lse), dP * m - D)
        lse_s[st * 128 + c] = qq < p.Lq ? -p.lse2[o] : -INFINITY;
        dv_s[st * 128 + c] = qq < p.Lq ? -p.dvec[o] : 0.f;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_stf + 8 * st);
    }
  } else if (warp == W_S) {
    // ---- S^T_h = K Q_h^T and dP^T_h = V dO_h^T of the NEXT query tile, as soon as warpgroup h has consumed the
    //      current ones
    if (lane == 0 && n_it > 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, false, false);
      const uint32_t bK = base_k64(sK), bV = base_k64(sV);
      const uint32_t bQk = base_k64(sQ), bdOk = base_k64(sdO);
      auto issue_s = [&](int hh, int st) {
        const uint32_t off = (uint32_t)(st * QKV_BYTES + hh * 4096) >> 4;
        const uint32_t qh = bQk + off, doh = bdOk + off;
#pragma unroll
        for (int k = 0; k < 6; ++k) tc_mma_bf16_lh(tmem + 64 * hh, step_k64(bK, k), kHi64, step_k64(qh, k), kHi64, idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < 6; ++k) tc_mma_bf16_lh(tmem + 128 + 64 * hh, step_k64(bV, k), kHi64, step_k64(doh, k), kHi64, idesc_s, k > 0);
        tc_commit(bar_sf + 8 * hh);
      };
      mbar_wait(bar_kv, 0);
      mbar_wait(bar_qf, 0);
      tc_fence_after();
      issue_s(0, 0);
      issue_s(1, 0);
      tc_commit(bar_qe);  // this issuer's last MMAs on stage 0
      for (int it = 0; it + 1 < n_it; ++it) {
        const int sn = (it + 1) & 1;
        mbar_wait(bar_qf + 8 * sn, ((it + 1) >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          mbar_wait(bar_sc + 8 * hh, it & 1);  // S^T_h / dP^T_h of tile `it` are in registers
          tc_fence_after();
          issue_s(hh, sn);
        }
        tc_commit(bar_qe + 8 * sn);
      }
    }
  } else if (warp == W_G) {
    // ---- dV += P^T_h dO_h (A = P^T_h packed bf16 in TMEM) and dK += dS^T_h Q_h (A = dS^T_h in shared memory, which
    //      is also what the workspace store reads)
    if (lane == 0 && n_it > 0) {
      constexpr uint32_t idesc_g = umma_idesc_bf16(128, DH, false, true);
      const uint32_t bQm = base_mn64(sQ), bdOm = base_mn64(sdO);  // MN-major views (B operands)
      for (int it = 0; it < n_it; ++it) {
        const int st = it & 1;
        mbar_wait(bar_qf + 8 * st, (it >> 1) & 1);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          mbar_wait(bar_pf + 8 * hh, it & 1);
          tc_fence_after();
          const uint32_t off = (uint32_t)(st * QKV_BYTES + hh * 4096) >> 4;
          const uint32_t qh = bQm + off, doh = bdOm + off;
          const uint32_t acc = (it > 0 || hh > 0) ? 1u : 0u;
          if (p.write_ds) {  // dS^T_h [128 keys x 64 queries] -> workspace [b*H + h][key][query] for the dQ kernel
            tma_store_3d(&tmDS, sDST + hh * HALF_BYTES, (i_begin + it) * TILE + hh * 64, kv0, b * p.H + h);
            tma_commit_group();
          }
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_bf16_ts(tmem + 256, tmem + 448 + 32 * hh + 8 * k, step_mn64(doh, k), kHi64, idesc_g, (acc || k > 0) ? 1u : 0u);
          const uint32_t bDS = umma_desc_lo(sDST + hh * HALF_BYTES, 16);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_bf16_lh(tmem + 352, bDS + 2 * k, kHi128, step_mn64(qh, k), kHi64, idesc_g, (acc || k > 0) ? 1u : 0u);
          if (hh == 1) tc_commit(bar_qe + 8 * st);  // this issuer's last MMAs on Q / dO stage `st`
          // P^T_h / dS^T_h go back to warpgroup hh once these MMAs have retired and the workspace store has read dS^T_h
          if (p.write_ds) tma_wait_group_read0();
          tc_commit(bar_gd + 8 * hh);
        }
      }
      tc_commit(bar_fin);  // every gradient MMA has retired: dV / dK may be drained (both warpgroups wait on this one;
                           // a parity wait on the other half's `gd` barrier can pass a phase too early)
      if (p.write_ds) tma_wait_group0();  // workspace writes complete before the grid ends
    }
  } else {
    // warps [0, AW / 2) own query columns [0, 64) of every tile (hh = 0), the others [64, 128); with AW = 16 two warps
    // of the same TMEM lane quadrant share a key row: sub = 0 takes the first 32 of the half's 64 queries, sub = 1 the rest
    constexpr int NSUB = AW / 8;
    const int hh = warp / (AW / 2), quad = warp & 3, sub = NSUB == 2 ? (warp >> 2) & 1 : 0;
    const int r = quad * 32 + lane;  // local key row
    const int kv = kv0 + r;
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tST = tmem + 64 * hh, tDPT = tmem + 128 + 64 * hh, tPT = tmem + 448 + 32 * hh;
    const uint32_t sDSTh = sDST + hh * HALF_BYTES;
    bool row_valid = kv < p.Lk;
    if (row_valid && p.kpm) row_valid = p.kpm[(long long)b * p.Lk + kv] == 0;
    const uint32_t bh = (uint32_t)(b * p.H + h);
    const DropKey dkey = drop_key(p);
    for (int it = 0; it < n_it; ++it) {
      const int st = it & 1;
      const int qi = i_begin + it;
      const int q0 = qi * TILE + hh * 64;  // first query of this half
      // keep bits of (q = q0 + c0 + i, k = this thread's key row): lane L makes the word of query q0 + c0 + L over the
      // warp's 32 keys, then the warp transposes the 32x32 bit tile (index-only work, done before the waits)
      uint32_t kw0 = keep_word(p, dkey, bh, (uint32_t)(q0 + 32 * sub + lane), (uint32_t)((kv0 >> 5) + quad));
      uint32_t kw1 = NSUB == 1 ? keep_word(p, dkey, bh, (uint32_t)(q0 + 32 + lane), (uint32_t)((kv0 >> 5) + quad)) : 0u;
      if (p.thresh16 != 0) {
        kw0 = warp_bit_transpose(kw0, lane);
        if (NSUB == 1) kw1 = warp_bit_transpose(kw1, lane);
      }
      mbar_wait(bar_stf + 8 * st, (it >> 1) & 1);
      mbar_wait(bar_sf + 8 * hh, it & 1);
      tc_fence_after();
      const bool diag = p.causal && (qi == jt);
      if (it > 0) mbar_wait(bar_gd + 8 * hh, (it - 1) & 1);  // P^T_h / dS^T_h buffers free again
      // warp-uniform variants: only the diagonal tile pays for the causal comparison, only warps that hold a masked /
      // out-of-range key row pay for the row select
      const float* nlse = lse_s + st * 128 + hh * 64;
      const float* ndv = dv_s + st * 128 + hh * 64;
      const bool rowmask = !__all_sync(0xffffffffu, row_valid);
      if (diag) {
        if (rowmask) dkdv_half_tile<true, true, NSUB>(tST, tDPT, tPT, lane_sel, sDSTh, nlse, ndv, kw0, kw1, r, hh, row_valid, p.scale_log2, p.inv_keep, sub, bar_sc + 8 * hh);
        else dkdv_half_tile<true, false, NSUB>(tST, tDPT, tPT, lane_sel, sDSTh, nlse, ndv, kw0, kw1, r, hh, row_valid, p.scale_log2, p.inv_keep, sub, bar_sc + 8 * hh);
      } else {
        if (rowmask) dkdv_half_tile<false, true, NSUB>(tST, tDPT, tPT, lane_sel, sDSTh, nlse, ndv, kw0, kw1, r, hh, row_valid, p.scale_log2, p.inv_keep, sub, bar_sc + 8 * hh);
        else dkdv_half_tile<false, false, NSUB>(tST, tDPT, tPT, lane_sel, sDSTh, nlse, ndv, kw0, kw1, r, hh, row_valid, p.scale_log2, p.inv_keep, sub, bar_sc + 8 * hh);
      }
      tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_pf + 8 * hh);
    }
    if (n_it > 0) {
      mbar_wait(bar_fin, 0);
      tc_fence_after();
    }
    {
      // warpgroup 0 drains dV (staged in the V tile's shared memory), warpgroup 1 dK * scale (in the K tile's): every
      // MMA has retired, the operand tiles are dead.  One thread per warpgroup then stores the tile with TMA (rows
      // past Lk are clipped by the tensor map).
      const uint32_t src = tmem + (hh == 0 ? 256 : 352);
      const float sc = hh == 0 ? 1.0f : p.scale;
      const uint32_t stage = hh == 0 ? sV : sK;
#pragma unroll 1
      for (int c = (NSUB == 2 && sub == 1) ? 2 : 0; c < ((NSUB == 2 && sub == 0) ? 2 : 3); ++c) {
        uint32_t rr[32];
        if (n_it > 0) {
          tmem_ld32(src + lane_sel + c * 32, rr);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) rr[i] = 0u;
        }
        stage_chunk_sw64(stage, r, c, rr, sc);
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + hh, 16 * AW);
      if (quad == 0 && lane == 0 && sub == 0) {
        store_tile(hh == 0 ? &tmDV : &tmDK, stage, h * DH, kv0, b);
        tma_commit_group();
        tma_wait_group_read0();  // shared memory stays valid until the store has read it
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_S) tmem_dealloc(tmem, 512);
}


// ---- dQ from stored dS^T: one CTA per (q-tile, head, batch), three co-resident per SM ----
// The dK/dV kernel already computed every dS^T tile (exp, dropout mask, dP - D) and left it in the workspace
// [B*H][Lk][Lq] (bf16), so dQ = scale * sum_j dS_j K_j is a plain TMA -> tcgen05 stream (no second softmax /
// dropout recomputation, which is what bounds the recomputing dQ kernel).  A pipeline stage holds 64 keys:
// dS^T [64 keys x 128 queries] as two swizzle-128 blocks (the MN-major A operand: M = queries contiguous) and
// K [64 keys x 96] as three swizzle-64 blocks (MN-major B operand).  Two stages per CTA, THREE CTAs per SM: the same
// six stages in flight per SM as 2 x 3, but a CTA's prologue (TMEM allocation, first loads) and epilogue (dQ drain +
// store) now overlap with two streaming neighbours instead of one (385 -> 380 us per backward with key-padding masks).
constexpr int DQ2_STAGES = 2;
constexpr int DQ2_A_BYTES = 2 * 64 * 128;   // two [64 keys x 64 queries] blocks
constexpr int DQ2_B_BYTES = 3 * 64 * 64;    // three [64 keys x 32 d] blocks
constexpr int DQ2_STAGE_BYTES = DQ2_A_BYTES + DQ2_B_BYTES;
constexpr int DQ2_SMEM = 1024 + DQ2_STAGES * DQ2_STAGE_BYTES + 256;
constexpr int DQ2_THREADS = 192;

__global__ void __launch_bounds__(DQ2_THREADS, 3)
attn_bwd_dq_ds_kernel(const __grid_constant__ CUtensorMap tmDS, const __grid_constant__ CUtensorMap tmK64,
                      const __grid_constant__ CUtensorMap tmDQ, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t aux = base + DQ2_STAGES * DQ2_STAGE_BYTES;
  const uint32_t bar_full = aux, bar_empty = aux + 32, bar_done = aux + 64, tmem_ptr_addr = aux + 72;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(gen + (tmem_ptr_addr - base));
  int* ext_slot = reinterpret_cast<int*>(gen + (aux + 80 - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nq_tiles = gridDim.x;
  const int qt = nq_tiles - 1 - blockIdx.x;
  // the dK/dV kernel wrote the workspace in (b, h) order: start with what it wrote LAST, which is still in L2
  const int h = gridDim.y - 1 - blockIdx.y, b = gridDim.z - 1 - blockIdx.z;
  const int q0 = qt * TILE;

  if (tid == 0) {
    for (int i = 0; i < DQ2_STAGES; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    mbar_init(bar_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_addr, 128);
  int nkv = (kv_extent(p, b, ext_slot) + TILE - 1) / TILE;
  if (p.causal) nkv = min(nkv, qt + 1);
  const int n_st = 2 * nkv;  // 64-key stages
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_gen;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < n_st; ++i) {
        mbar_wait(bar_empty + 8 * s, ph ^ 1);
        const uint32_t full = bar_full + 8 * s;
        const uint32_t sA = base + s * DQ2_STAGE_BYTES, sB = sA + DQ2_A_BYTES;
        mbar_expect_tx(full, DQ2_STAGE_BYTES);
        tma_load_3d(&tmDS, full, sA, q0, i * 64, b * p.H + h);
        tma_load_3d(&tmDS, full, sA + 64 * 128, q0 + 64, i * 64, b * p.H + h);
#pragma unroll
        for (int c = 0; c < 3; ++c) tma_load_3d(&tmK64, full, sB + c * 4096, h * DH + 32 * c, i * 64, b);
        if (++s == DQ2_STAGES) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_st > 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, DH, true, true);
      constexpr uint32_t kHiB = umma_desc_hi(512, UMMA_SW64);
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < n_st; ++i) {
        mbar_wait(bar_full + 8 * s, ph);
        tc_fence_after();
        const uint32_t sA = base + s * DQ2_STAGE_BYTES, sB = sA + DQ2_A_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k)  // 16 keys per step: 2048 B of A rows, 1024 B of B rows
          tc_mma_bf16_lh(tmem, umma_desc_lo(sA + k * 2048, 64 * 128), kHi128, umma_desc_lo(sB + k * 1024, 4096), kHiB,
                         idesc, (i > 0 || k > 0) ? 1u : 0u);
        tc_commit(bar_empty + 8 * s);
        if (++s == DQ2_STAGES) {
          s = 0;
          ph ^= 1;
        }
      }
      tc_commit(bar_done);
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    if (n_st > 0) {
      mbar_wait(bar_done, 0);
      tc_fence_after();
    }
    // dQ * scale -> bf16, staged in the (dead) first pipeline stage and stored by one thread with TMA
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
      uint32_t rr[32];
      if (n_st > 0) {
        tmem_ld32(tmem + lane_sel + c * 32, rr);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) rr[i] = 0u;
      }
      stage_chunk_sw64(base, r, c, rr, p.scale);
    }
    fence_proxy_async_smem();
    named_bar_sync(1, 128);
    if (warp == 2 && lane == 0) {
      store_tile(&tmDQ, base, h * DH, q0, b);
      tma_commit_group();
      tma_wait_group_read0();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 128);
}

// -----------------------------------------------------------------------------------------------
int make_qkv_map(CUtensorMap* m, const void* ptr, int64_t ld, int64_t B, int64_t L, int64_t H,
                 int64_t batch_stride = 0) {
  const uint64_t bs = batch_stride > 0 ? (uint64_t)batch_stride : (uint64_t)L * ld;
  return make_tmap_3d(m, ptr, 2, (uint64_t)(H * DH), (uint64_t)L, (uint64_t)B, (uint64_t)ld * 2, bs * 2, 32, TILE,
                      SWZ_64);
}

int fill_params(AttnParams& p, int64_t B, int64_t H, int64_t Lq, int64_t Lk, int32_t causal, float scale,
                const uint8_t* kpm, float p_drop, uint64_t seed, uint64_t offset, const uint64_t* epoch) {
  SCT_CHECK(B > 0 && H > 0 && Lq > 0 && Lk > 0, "empty attention problem");
  SCT_CHECK(B <= 65535 && H <= 65535, "grid overflow");
  SCT_CHECK(!causal || Lq == Lk, "causal attention requires Lq == Lk");
  SCT_CHECK(p_drop >= 0.f && p_drop < 1.f, "p_drop out of range");
  p.B = (int)B; p.H = (int)H; p.Lq = (int)Lq; p.Lk = (int)Lk;
  p.causal = causal;
  p.scale = scale;
  p.scale_log2 = scale * kLog2e;
  p.kpm = kpm;
  {
    // stream key: splitmix64 of (seed, offset)
    uint64_t z = seed * 0x9E3779B97F4A7C15ull + offset * 0xD1B54A32D192ED03ull + 0x2545F4914F6CDD1Dull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    p.key0 = (uint32_t)z;
    p.key1 = (uint32_t)(z >> 32);
  }
  p.epoch = reinterpret_cast<const unsigned long long*>(epoch);
  const double full = (double)(1u << kDropBits);
  double t = (double)p_drop * full + 0.5;
  if (t < 1.0) t = 1.0;  // p_drop > 0 drops at least 2^-kDropBits
  p.thresh16 = p_drop > 0.f ? (uint32_t)(t > full - 1.0 ? full - 1.0 : t) : 0u;
  for (int i = 0; i < 16; ++i) p.tmask[i] = ((p.thresh16 >> i) & 1u) ? 0xFFFFFFFFu : 0u;
  // the scale uses the realised keep probability (thresh / 2^kDropBits is p_drop to within 2^-(kDropBits+1))
  p.inv_keep = p_drop > 0.f ? (float)(full / (full - (double)p.thresh16)) : 1.0f;
  p.write_ds = 0;
  p.lse2 = nullptr; p.dvec = nullptr; p.o = nullptr; p.dq = p.dk = p.dv = nullptr;
  p.ldo = p.ldq_out = p.ldkv_out = 0;
  return 0;
}

__global__ void read_flag_kernel(int* out) { *out = g_timeout_flag; }

}  // namespace

int attn_timeout_flag() {
  int* d = nullptr;
  int h = 0;
  if (cudaMalloc(&d, sizeof(int)) != cudaSuccess) return -1;
  read_flag_kernel<<<1, 1>>>(d);
  cudaMemcpy(&h, d, sizeof(int), cudaMemcpyDeviceToHost);
  cudaFree(d);
  return h;
}

}  // namespace sct

using namespace sct;

extern "C" {

int32_t sct_attn_fwd_strided(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv,
                             int64_t kv_batch_stride, void* o, int64_t ldo, float* lse2, const uint8_t* kpm,
                             int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t head_dim, int32_t causal,
                             float scale, float p_drop, uint64_t seed, uint64_t offset, const uint64_t* epoch, void* stream) {
  SCT_CHECK(q && k && v && o, "null pointer");
  SCT_CHECK(head_dim == DH, "head_dim %lld unsupported (kernel is specialised for 96)", (long long)head_dim);
  SCT_CHECK(ldo % 8 == 0, "ldo must be a multiple of 8");
  AttnParams p;
  if (int rc = fill_params(p, B, H, Lq, Lk, causal, scale, kpm, p_drop, seed, offset, epoch)) return rc;
  p.o = (__nv_bfloat16*)o;
  p.ldo = (int)ldo;
  p.lse2 = lse2;
  SCT_CHECK(kv_batch_stride == 0 || (kv_batch_stride >= Lk * ldkv && kv_batch_stride % 8 == 0),
            "kv_batch_stride must be 0 or a multiple of 8 >= Lk * ldkv");
  const int use_dec = env_int("SCT_ATTN_DECODE", 1);  // =0: tile kernel also for Lq = 1 (A/B timing)
  if (use_dec && Lq == 1 && p_drop == 0.f) {  // decode step: SIMT stream over the KV cache
    SCT_CHECK(ldq % 8 == 0 && ldkv % 8 == 0, "row pitches must be multiples of 8");
    const long long kv_bs = kv_batch_stride > 0 ? kv_batch_stride : Lk * ldkv;
    // more key groups per block for longer caches (measured at B = 128, H = 8: 128 threads win up to ~1024 cached rows)
    const int dec_threads = env_int("SCT_ATTN_DECODE_THREADS", -1);
    const int nt = dec_threads > 0 ? dec_threads : (Lk <= 1024 ? 128 : 256);
    const dim3 grid((unsigned)H, (unsigned)B);
    if (nt == 128)
      attn_decode_kernel<128><<<grid, 128, 0, (cudaStream_t)stream>>>(
          (const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, ldkv, kv_bs,
          (__nv_bfloat16*)o, ldo, lse2, kpm, (int)H, (int)Lk, p.scale_log2);
    else
      attn_decode_kernel<256><<<grid, 256, 0, (cudaStream_t)stream>>>(
          (const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, ldkv, kv_bs,
          (__nv_bfloat16*)o, ldo, lse2, kpm, (int)H, (int)Lk, p.scale_log2);
    SCT_LAUNCH_CHECK();
    return 0;
  }
  CUtensorMap tq, tk, tv, to;
  if (int rc = make_qkv_map(&tq, q, ldq, B, Lq, H)) return rc;
  if (int rc = make_qkv_map(&to, o, ldo, B, Lq, H)) return rc;
  if (int rc = make_qkv_map(&tk, k, ldkv, B, Lk, H, kv_batch_stride)) return rc;
  if (int rc = make_qkv_map(&tv, v, ldkv, B, Lk, H, kv_batch_stride)) return rc;
  dim3 grid((unsigned)((Lq + TILE - 1) / TILE), (unsigned)H, (unsigned)B);
  if (env_int("SCT_ATTN_FWD", 2) == 2 && Lq > 1) {  // two 64-key halves pipelined inside the CTA
    CUtensorMap tk64, tv64;
    const uint64_t bs = kv_batch_stride > 0 ? (uint64_t)kv_batch_stride : (uint64_t)Lk * ldkv;
    if (int rc = make_tmap_3d(&tk64, k, 2, (uint64_t)(H * DH), (uint64_t)Lk, (uint64_t)B, (uint64_t)ldkv * 2, bs * 2, 32, 64,
                              SWZ_64))
      return rc;
    if (int rc = make_tmap_3d(&tv64, v, 2, (uint64_t)(H * DH), (uint64_t)Lk, (uint64_t)B, (uint64_t)ldkv * 2, bs * 2, 32, 64,
                              SWZ_64))
      return rc;
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(attn_fwd2_kernel), FWD2_SMEM)) return rc;
    attn_fwd2_kernel<<<grid, FWD2_THREADS, FWD2_SMEM, (cudaStream_t)stream>>>(tq, tk64, tv64, to, p);
    SCT_LAUNCH_CHECK();
    return 0;
  }
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(attn_fwd_kernel), FWD_SMEM)) return rc;
  attn_fwd_kernel<<<grid, 256, FWD_SMEM, (cudaStream_t)stream>>>(tq, tk, tv, to, p);
  SCT_LAUNCH_CHECK();
  return 0;
}

int32_t sct_attn_fwd(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, void* o,
                     int64_t ldo, float* lse2, const uint8_t* kpm, int64_t B, int64_t H, int64_t Lq,
                     int64_t Lk, int64_t head_dim, int32_t causal, float scale, float p_drop,
                     uint64_t seed, uint64_t offset, const uint64_t* epoch, void* stream) {
  return sct_attn_fwd_strided(q, ldq, k, v, ldkv, 0, o, ldo, lse2, kpm, B, H, Lq, Lk, head_dim, causal, scale, p_drop,
                              seed, offset, epoch, stream);
}

int32_t sct_attn_bwd(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv,
                     const void* o, const void* d_o, int64_t ldo, const float* lse2, float* dvec,
                     void* dq, int64_t lddq, void* dk, void* dv, int64_t lddkv, const uint8_t* kpm,
                     int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t head_dim, int32_t causal,
                     float scale, float p_drop, uint64_t seed, uint64_t offset, const uint64_t* epoch, void* stream) {
  return sct_attn_bwd_ws(q, ldq, k, v, ldkv, o, d_o, ldo, lse2, dvec, dq, lddq, dk, dv, lddkv, kpm, B, H, Lq, Lk,
                         head_dim, causal, scale, p_drop, seed, offset, epoch, nullptr, 0, stream);
}

int64_t sct_attn_bwd_workspace_bytes(int64_t B, int64_t H, int64_t Lq, int64_t Lk) {
  const int64_t pitch = (Lq + 63) / 64 * 64;
  return B * H * Lk * pitch * 2;
}

int32_t sct_attn_bwd_ws(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv,
                        const void* o, const void* d_o, int64_t ldo, const float* lse2, float* dvec,
                        void* dq, int64_t lddq, void* dk, void* dv, int64_t lddkv, const uint8_t* kpm,
                        int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t head_dim, int32_t causal,
                        float scale, float p_drop, uint64_t seed, uint64_t offset, const uint64_t* epoch, void* workspace,
                        int64_t workspace_bytes, void* stream) {
  SCT_CHECK(q && k && v && o && d_o && lse2 && dvec && dq && dk && dv, "null pointer");
  const bool use_ws = workspace != nullptr;
  SCT_CHECK(!use_ws || workspace_bytes >= sct_attn_bwd_workspace_bytes(B, H, Lq, Lk),
            "attention backward workspace too small (%lld < %lld bytes)", (long long)workspace_bytes,
            (long long)sct_attn_bwd_workspace_bytes(B, H, Lq, Lk));
  SCT_CHECK(head_dim == DH, "head_dim %lld unsupported (kernel is specialised for 96)", (long long)head_dim);
  SCT_CHECK(H * DH == 768 || H * DH <= ldo, "unexpected head layout");
  SCT_CHECK(lddq % 8 == 0 && lddkv % 8 == 0 && ldo % 8 == 0, "gradient pitches must be multiples of 8");
  AttnParams p;
  if (int rc = fill_params(p, B, H, Lq, Lk, causal, scale, kpm, p_drop, seed, offset, epoch)) return rc;
  p.lse2 = const_cast<float*>(lse2);
  p.dvec = dvec;
  p.dq = (__nv_bfloat16*)dq; p.dk = (__nv_bfloat16*)dk; p.dv = (__nv_bfloat16*)dv;
  p.ldq_out = (int)lddq; p.ldkv_out = (int)lddkv; p.ldo = (int)ldo;
  cudaStream_t st = (cudaStream_t)stream;
  SCT_CHECK(H * DH == 32 * 24, "D-vector kernel assumes 8 heads x 96 (row width 768)");
  {
    const int rows = (int)(B * Lq);
    attn_bwd_dvec_kernel<<<(rows + 7) / 8, 256, 0, st>>>((const __nv_bfloat16*)o, (const __nv_bfloat16*)d_o,
                                                         dvec, rows, (int)Lq, (int)H, (int)ldo);
    SCT_LAUNCH_CHECK();
  }
  CUtensorMap tq, tk, tv, tdo, tds_st, tds_ld, tk64;
  if (int rc = make_qkv_map(&tq, q, ldq, B, Lq, H)) return rc;
  if (int rc = make_qkv_map(&tk, k, ldkv, B, Lk, H)) return rc;
  if (int rc = make_qkv_map(&tv, v, ldkv, B, Lk, H)) return rc;
  if (int rc = make_qkv_map(&tdo, d_o, ldo, B, Lq, H)) return rc;
  tds_st = tq;  // placeholder when the workspace is not used (never dereferenced)
  if (use_ws) {
    // workspace = dS^T [B*H][Lk][pitch >= Lq] bf16; stored as [128 keys x 64 queries] blocks, re-read 64 keys at a time
    const uint64_t pitch = (uint64_t)((Lq + 63) / 64 * 64);
    if (int rc = make_tmap_3d(&tds_st, workspace, 2, (uint64_t)Lq, (uint64_t)Lk, (uint64_t)(B * H), pitch * 2,
                              (uint64_t)Lk * pitch * 2, 64, TILE, SWZ_128))
      return rc;
    if (int rc = make_tmap_3d(&tds_ld, workspace, 2, (uint64_t)Lq, (uint64_t)Lk, (uint64_t)(B * H), pitch * 2,
                              (uint64_t)Lk * pitch * 2, 64, 64, SWZ_128))
      return rc;
    if (int rc = make_tmap_3d(&tk64, k, 2, (uint64_t)(H * DH), (uint64_t)Lk, (uint64_t)B, (uint64_t)ldkv * 2,
                              (uint64_t)Lk * ldkv * 2, 32, 64, SWZ_64))
      return rc;
    p.write_ds = 1;
  }
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(attn_bwd_dkdv_kernel<8>), BWD3_KV_SMEM)) return rc;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(attn_bwd_dkdv_kernel<16>), BWD3_KV_SMEM)) return rc;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(attn_bwd_dq_kernel), BWD3_Q_SMEM)) return rc;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(attn_bwd_dq_ds_kernel), DQ2_SMEM)) return rc;
  {
    dim3 grid((unsigned)((Lk + TILE - 1) / TILE), (unsigned)H, (unsigned)B);
    CUtensorMap tdk, tdv;
    if (int rc = make_qkv_map(&tdk, dk, lddkv, B, Lk, H)) return rc;
    if (int rc = make_qkv_map(&tdv, dv, lddkv, B, Lk, H)) return rc;
    if (env_int("SCT_ATTN_BWD_WARPS", 8) == 16)  // (measured equal or slower than 8: the arithmetic warps are not the limiter)
      attn_bwd_dkdv_kernel<16><<<grid, bwd_kv_threads(16), BWD3_KV_SMEM, st>>>(tq, tk, tv, tdo, tds_st, tdk, tdv, p);
    else
      attn_bwd_dkdv_kernel<8><<<grid, bwd_kv_threads(8), BWD3_KV_SMEM, st>>>(tq, tk, tv, tdo, tds_st, tdk, tdv, p);
    SCT_LAUNCH_CHECK();
  }
  {
    dim3 grid((unsigned)((Lq + TILE - 1) / TILE), (unsigned)H, (unsigned)B);
    if (use_ws) {
      CUtensorMap tdq;
      if (int rc = make_qkv_map(&tdq, dq, lddq, B, Lq, H)) return rc;
      attn_bwd_dq_ds_kernel<<<grid, DQ2_THREADS, DQ2_SMEM, st>>>(tds_ld, tk64, tdq, p);
    } else {
      attn_bwd_dq_kernel<<<grid, BWD3_THREADS, BWD3_Q_SMEM, st>>>(tq, tk, tv, tdo, p);
    }
    SCT_LAUNCH_CHECK();
  }
  return 0;
}

}  // extern "C"
