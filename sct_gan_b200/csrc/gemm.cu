// sct_b200 — K2: persistent, warp-specialised bf16 GEMM on tcgen05 / TMEM, fed by TMA.
//
// Replaces every nn.Linear / MHA projection / FFN / vocab-projection matmul of the reference hot path
// (SCT-GAN/model.py:56-82, 209-271; torch nn/functional.py linear -> cuBLASLt addmm) and their
// autograd backward (dgrad, wgrad).
//
//   D[M,N] = A[M,K] * B[N,K]^T (+ bias[N])
//
// Each operand may be K-major (the contraction index is contiguous in global memory) or MN-major (the
// M / N index is contiguous).  All three training GEMMs are therefore served by one kernel without any
// transposed copy in HBM:
//   forward  Y  = X  W^T : A = X  [M,K]   K-major,  B = W  [N,K]   K-major
//   dgrad    dX = dY W   : A = dY [M,K'] K-major,   B = W  [K',N'] MN-major (same bytes as forward W)
//   wgrad    dW = dY^T X : A = dY [K',M'] MN-major, B = X  [K',N'] MN-major, fp32 split-K reduce-add
//
// Layout in shared memory (per pipeline stage): operand tiles are stored as column blocks of
// [rows x 128 B] with the TMA/UMMA 128-byte swizzle; a K-major tile is one block [128|BN rows x 64 k],
// an MN-major tile is BM/64 (BN/64) blocks of [64 k-rows x 64 mn].
//
// Roles (256 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2-5 = epilogue (TMEM -> registers -> swizzled smem -> TMA store / TMA reduce-add), warps 6-7 = optional
// column sums of the A operand (wgrad only: the bias gradient db = colsum(dY) read from the dY tiles that are in
// shared memory anyway, instead of a separate pass over dY in HBM).
// Two TMEM accumulator stages let the epilogue of tile i overlap the mainloop of tile i+1.
//
// CLUSTER = 2 (cta_group::2): the two CTAs of a cluster (the two SMs of a TPC) compute ONE 256 x BN tile with
// 256-row UMMAs issued by the leader CTA's MMA thread.  Each CTA loads its own 128 rows of A and only HALF of the
// B tile (BN / 2 rows): the tensor cores of both SMs read both halves, so a pipeline stage is 32 KB instead of
// 48 KB per SM (6 stages instead of 4, a third less L2 -> SM traffic, half the MMA issue work per SM).  Every
// TMA load of the pair is counted on the LEADER's `full` barrier; tcgen05.commit multicasts the stage release
// (`empty`) and the accumulator hand-off (`tfull`) to both CTAs; both epilogues arrive on the leader's `tempty`.
#include <stdlib.h>

#include <atomic>
#include <mutex>

#include "../../include/sct_b200.h"
#include "common.cuh"

namespace sct {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kThreads = 256;      // plain kernel: TMA, MMA, 4 epilogue warps, 2 column-sum warps (wgrad)
// fused-epilogue kernels: TMA, MMA + epilogue warps.  EPI_MUL: 8 (two per TMEM lane quadrant); EPI_GELU_FWD: 16 (four
// per quadrant: the activation is a long dependent instruction stream per element, and a scheduler with only two such
// warps issues ~0.3 instructions per clock)
__host__ __device__ constexpr int epi_warps(int epi) { return epi == 1 /*EPI_GELU_FWD*/ ? 16 : 8; }
__host__ __device__ constexpr int epi_threads(int epi) { return 64 + 32 * epi_warps(epi); }
constexpr int kSmemLimit = 232448;  // 227 KB

// Fused epilogues (SURVEY §8b epilogue enum) for the feed-forward block.
//   EPI_GELU_FWD: the first linear also applies the activation.  z = A W^T + b stays in registers; the kernel stores
//     H = dropout(gelu(z)) (what linear2 consumes) and G = mask / (1 - p) * gelu'(z), the LOCAL DERIVATIVE of that
//     activation + dropout, instead of z itself (torch transformer.py _ff_block: linear1 -> activation -> dropout).
//     The separate gelu_dropout pass over [M, ff] disappears and nothing has to be recomputed in backward.
//   EPI_MUL: the dgrad GEMM of the second linear turns dH = dY W2 into dZ = dH * G on the way out (one multiply per
//     element, no transcendental, no mask regeneration).
enum Epi { EPI_NONE = 0, EPI_GELU_FWD = 1, EPI_MUL = 2 };

constexpr int kEpiDropBits = 8;  // drop probability resolved to 2^-8, as in the attention kernels
struct EpiParams {
  const __nv_bfloat16* z;  // EPI_MUL: elementwise multiplier G [M, ldz]
  long long ldz;
  uint32_t k0, k1;         // dropout stream key (seed, offset mixed on the host)
  const unsigned long long* epoch;  // device-resident epoch folded into the key at run time (nullable)
  uint32_t tmask[kEpiDropBits];     // bit i of the threshold spread over a word
  uint32_t thresh;         // 0 = no dropout
  float inv_keep;
};

struct GemmParams {
  EpiParams epi;
  int* sched;  // [2] device ints, zero on entry and on exit: next work item, leaders that have drained the queue
  int dynamic; // 1: items drawn from the counter; 0: static stride over the grid (SCT_GEMM_DYNAMIC=0, A/B timing)
  int M, N, K;
  int m_tiles, n_tiles, k_splits;
  int kb_total, kb_per_split;
  int n_fast;         // 1: consecutive tiles walk N first (A tile reused from L2), 0: walk M first
  const float* bias;  // nullable, fp32 [N]
  float alpha;        // output scale applied before bias
  float* colsum;      // nullable (MN-major A only), fp32 [M]: += alpha * sum_k A[k, m]
};

template <int BN, bool OUT_F32, int CLUSTER, int EPI = EPI_NONE>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN / CLUSTER * BK * 2;  // this CTA's share of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // Epilogue staging: two [128 rows x 128 B] blocks (64 bf16 / 32 fp32 columns each), ping-ponged chunk by chunk.
  // A full-tile staging buffer (64 KB at BN = 256) would leave only 3 pipeline stages, and the mainloop is
  // bound by the bytes it can keep in flight (measured: the MMA thread waited on `full` 44 % of the time).
  // Fused epilogues use four blocks: EPI_GELU_FWD stores two tensors per chunk (H and G) for each of its two warp
  // groups; EPI_MUL receives the multiplier chunk by TMA in the block it later stores the product from.
  static constexpr int C_BLOCKS = EPI != EPI_NONE ? 4 : 2;
  static constexpr int C_BYTES = C_BLOCKS * BM * 128;
  static constexpr int AUX_BYTES = 1024;  // barriers + tmem ptr
  static constexpr int BIAS_BYTES = (EPI != EPI_NONE ? 2 : 1) * BN * 4;  // fused epilogues: one copy per accumulator stage
  static constexpr int STAGES_RAW = (kSmemLimit - 1024 - C_BYTES - AUX_BYTES - BIAS_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + C_BYTES + AUX_BYTES + BIAS_BYTES;
  static constexpr int THREADS = EPI != EPI_NONE ? epi_threads(EPI) : kThreads;
  static constexpr int TMEM_COLS = 2 * BN;  // two accumulator stages (256 or 512, powers of two)
  static_assert(STAGES >= 2, "pipeline too shallow");
};

struct TileCoord {
  int m_blk, n_blk, kb0, kb1;
};

__device__ __forceinline__ TileCoord decode_tile(const GemmParams& p, int t) {
  TileCoord c;
  const int ks = t % p.k_splits;
  const int t2 = t / p.k_splits;
  if (p.n_fast) {
    c.n_blk = t2 % p.n_tiles;
    c.m_blk = t2 / p.n_tiles;
  } else {
    c.m_blk = t2 % p.m_tiles;
    c.n_blk = t2 / p.m_tiles;
  }
  c.kb0 = ks * p.kb_per_split;
  c.kb1 = min(c.kb0 + p.kb_per_split, p.kb_total);
  return c;
}

// ---- fused-epilogue arithmetic --------------------------------------------------------------------------------
// Exact-erf GELU (activation='gelu') to bf16 accuracy with ONE transcendental per element:
//   Phi(x) = 0.5 (1 + erf(x / sqrt 2))  ~  0.5 (1 + tanh(x (a + b x^2)))      (a, b fitted against scipy's erf)
// max |x Phi_fit - x Phi| = 2.7e-4 and max |d/dx| error = 8.7e-4 over the real line — below the bf16 rounding of the
// stored results (half an ulp is 2e-3 at 1) — and tanh.approx.f32 is a single MUFU.  The derivative is the exact
// derivative of the fitted function: gelu'(x) = Phi + 0.5 x (1 - t^2)(a + 3 b x^2), so forward and backward are a
// consistent pair.  All multiplies / FMAs run as packed fp32 pairs (FFMA2 / FMUL2): ~11 instructions per element for
// BOTH outputs, against ~15 + two MUFUs for the Abramowitz-Stegun erf of the stand-alone kernels — an epilogue warp
// group has to finish its share of the 128 x 256 tile inside the mainloop of the next one.
constexpr float kGa = 0.80015708f, kGb = 0.03470089f;
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// (gelu(v), gelu'(v)) for a pair of pre-activations
__device__ __forceinline__ void gelu_pair(float2 v, float2& val, float2& grad) {
  const float2 x2 = __fmul2_rn(v, v);
  const float2 in = __fmul2_rn(v, __ffma2_rn(x2, make_float2(kGb, kGb), make_float2(kGa, kGa)));
  const float2 t = make_float2(tanh_approx(in.x), tanh_approx(in.y));
  const float2 phi = __ffma2_rn(t, make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f));
  val = __fmul2_rn(v, phi);
  // -0.5 x (a + 3 b x^2) and t^2 - 1: gelu' = Phi + [0.5 x (a + 3 b x^2)] (1 - t^2)
  const float2 tn = __fmul2_rn(v, __ffma2_rn(x2, make_float2(-1.5f * kGb, -1.5f * kGb), make_float2(-0.5f * kGa, -0.5f * kGa)));
  const float2 w = __ffma2_rn(t, t, make_float2(-1.f, -1.f));
  grad = __ffma2_rn(tn, w, phi);
}
__device__ __forceinline__ uint32_t epi_fmix32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}
// 32 keep bits for columns [32 cb, 32 cb + 32) of row `row`: bit-sliced threshold comparison of kEpiDropBits-bit
// uniforms (one strong hash per word, then one xorshift-multiply + one LOP3 per bit), ~1.5 instructions per element.
// The forward and the backward epilogue own the same (row, 32-column block) and regenerate the identical word.
__device__ __forceinline__ uint32_t epi_keep_word(const EpiParams& e, uint32_t k0, uint32_t k1, uint32_t row, uint32_t cb) {
  if (e.thresh == 0) return 0xFFFFFFFFu;
  uint32_t x = epi_fmix32((row * 0x9E3779B1u) ^ k0);
  x = epi_fmix32(x ^ (cb * 0x85EBCA77u) ^ k1);
  uint32_t lt = 0u;
#pragma unroll
  for (int i = 0; i < kEpiDropBits; ++i) {
    x = (x ^ (x >> 15)) * 0x2C1B3C6Du;
    const uint32_t nw = ~x, tm = e.tmask[i];
    lt = (nw & lt) | (tm & (nw | lt));
  }
  return ~lt;
}

// ---- dynamic work distribution ---------------------------------------------------------------------------------
// Work items (tiles, or pairs of tiles) are handed out by an atomic counter instead of a static stride over the grid:
// a persistent kernel that ASSUMES all 148 SMs loses a whole extra wave when some are busy with something else — the
// NCCL all-reduce kernels of the data-parallel gradient exchange hold 16-32 SMs for ~0.3 ms per bucket while backward
// keeps launching GEMMs whose 227 KB CTAs cannot share an SM with them (measured: +1.3 ms per step at 2 GPUs).  With
// the queue, CTAs that start late simply find less (or nothing) left.
// The leader CTA's producer thread draws the item and publishes it through a 2-slot queue in shared memory (of both
// CTAs of a pair: DSMEM store + remote arrive); every role of the CTA(s) — MMA issuer, epilogue warps, column-sum
// warps, the peer's producer — consumes the slot and hands it back on the LEADER's `empty` barrier.
// Cross-CTA signalling uses RELAXED cluster-scope operations only: a `.release.cluster` arrive compiles to
// MEMBAR.ALL.GPU (it has to drain the thread's outstanding global traffic — the prefetched atomic, in-flight stores),
// about a microsecond on the producer's critical path per tile (measured: -11 % throughput).  The leader therefore
// sends the item to the peer CTA as ONE tagged word (sequence number | item) that the peer's roles poll for in their
// own shared memory: a single-word message needs no ordering with anything else.
struct TileQueue {
  uint32_t full;          // leader CTA: its full[2] barriers
  uint32_t empty_leader;  // the leader CTA's empty[2] barriers (shared::cluster address when remote)
  const volatile uint32_t* slots;
  int qs;
  uint32_t ph;
  uint32_t seq;           // items consumed so far (tag expected in the slot)
  bool remote;
  // consumer: all lanes of the calling warp (or a single thread) get the item; one lane releases the slot
  __device__ __forceinline__ int next(bool whole_warp) {
    uint32_t m;
    if (!remote) {
      mbar_wait(full + 8 * qs, ph);
      m = slots[qs];
    } else {
      uint64_t t0 = 0;
#pragma unroll 1
      for (uint32_t spins = 1;; ++spins) {
        m = slots[qs];
        if ((m >> 24) == (seq & 0xFFu)) break;
        __nanosleep(32);
        if ((spins & 0x3FFFu) == 0 && mbar_wait_timeout_check(t0)) return -1;  // debug safety net: give up, drain
      }
    }
    if (whole_warp) __syncwarp();
    if (!whole_warp || (threadIdx.x & 31) == 0) {
      if (remote)
        mbar_arrive_cluster_relaxed(empty_leader + 8 * qs);
      else
        mbar_arrive(empty_leader + 8 * qs);
    }
    if ((qs ^= 1) == 0) ph ^= 1;
    ++seq;
    return (int)(m & 0xFFFFFFu) - 1;
  }
};
__device__ __forceinline__ uint32_t tq_message(uint32_t seq, int t) { return ((seq & 0xFFu) << 24) | (uint32_t)(t + 1); }

template <int BN, bool A_MN, bool B_MN, bool OUT_F32, int CLUSTER, int EPI>
__global__ void __launch_bounds__(EPI != EPI_NONE ? epi_threads(EPI) : kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmD2, const GemmParams p) {
  using C = Cfg<BN, OUT_F32, CLUSTER, EPI>;
  constexpr int STAGES = C::STAGES;
  constexpr bool PAIR = CLUSTER == 2;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sC = smem_base + STAGES * C::STAGE_BYTES;
  const uint32_t sAux = sC + C::C_BYTES;
  // aux: full[STAGES] | empty[STAGES] | tfull[2] | tempty[2] | tmem_ptr
  const uint32_t bar_full = sAux;
  const uint32_t bar_empty = sAux + 8 * STAGES;
  const uint32_t bar_tfull = sAux + 16 * STAGES;
  const uint32_t bar_tempty = bar_tfull + 16;
  const uint32_t tmem_ptr_addr = bar_tempty + 16;
  const uint32_t bar_done = sAux + 512;  // [STAGES] colsum mode: the MMAs reading this stage have retired
  const uint32_t bar_g = sAux + 640;     // [4] EPI_MUL: the multiplier chunk has landed in staging block i
  const uint32_t tq_full = sAux + 704, tq_empty = sAux + 720;  // [2] + [2] work-item queue
  const uint32_t tq_slots = sAux + 736;                        // [2] ints
  const volatile uint32_t* tq_gen = reinterpret_cast<const volatile uint32_t*>(smem_gen + (tq_slots - smem_base));
  volatile uint32_t* tmem_ptr_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_addr - smem_base));
  float* bias_s = reinterpret_cast<float*>(smem_gen + (sAux + C::AUX_BYTES - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // work items: with CLUSTER = 2 an item is a PAIR of vertically adjacent tiles (p.m_tiles then counts pairs)
  const int total_tiles = p.m_tiles * p.n_tiles * p.k_splits;
  const int crank = CLUSTER > 1 ? (int)cluster_ctarank() : 0;
  const int n_leaders = CLUSTER > 1 ? (int)(gridDim.x / CLUSTER) : (int)gridDim.x;
  const bool cs_active = A_MN && EPI == EPI_NONE && p.colsum != nullptr;  // column-sum warps take part in the queue
  constexpr int kEpiWarps = EPI != EPI_NONE ? epi_warps(EPI) : 4;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmD);
    if (EPI != EPI_NONE) tma_prefetch_desc(&tmD2);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
      mbar_init(bar_done + 8 * s, 1);
    }
    for (int s = 0; s < 4; ++s) mbar_init(bar_g + 8 * s, 1);
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(tq_slots), "r"(0x7F000000u) : "memory");      // tags no consumer expects
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(tq_slots + 4), "r"(0x7F000000u) : "memory");
    {
      // consumers of a queue slot: leader CTA = MMA thread + epilogue warps (+ 2 column-sum warps); peer CTA = its
      // producer thread + epilogue warps (+ column-sum warps).  All of them arrive on the LEADER's `empty`.
      const int per_cta = kEpiWarps + (cs_active ? 2 : 0) + 1;
      for (int s = 0; s < 2; ++s) {
        mbar_init(tq_full + 8 * s, 1);
        mbar_init(tq_empty + 8 * s, per_cta * CLUSTER);
      }
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tempty + 8 * s, kEpiWarps * CLUSTER);  // one arrival per epilogue warp (of both CTAs of a pair)
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_pair(tmem_ptr_addr, C::TMEM_COLS);
    else tmem_alloc(tmem_ptr_addr, C::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // the peer's barriers are initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  TileQueue tq;
  tq.full = tq_full;
  tq.remote = CLUSTER > 1 && crank != 0;
  tq.empty_leader = tq.remote ? cluster_map(tq_empty, 0) : tq_empty;
  tq.slots = tq_gen;
  tq.qs = 0;
  tq.ph = 0;
  tq.seq = 0;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      int gq = 0;           // leader: queue slot to fill next
      uint32_t gph = 0, gseq = 0;
      // the draw for item i + 1 is issued while the loads of item i are being queued: the ~1 us round trip of the
      // atomic would otherwise sit on the producer's critical path once per tile (measured: -10 % GEMM throughput)
      const int leader_id = CLUSTER > 1 ? (int)(blockIdx.x / CLUSTER) : (int)blockIdx.x;
      int t_next = crank == 0 ? (p.dynamic ? atomicAdd(p.sched, 1) : leader_id) : 0;
      while (true) {
        int t;
        if (crank == 0) {   // publish the drawn work item to every role of the CTA (pair)
          t = t_next < total_tiles ? t_next : -1;
          if (t >= 0) t_next = p.dynamic ? atomicAdd(p.sched, 1) : t_next + n_leaders;
          mbar_wait(tq_empty + 8 * gq, gph ^ 1);
          const uint32_t msg = tq_message(gseq++, t);
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(tq_slots + 4 * gq), "r"(msg) : "memory");
          mbar_arrive(tq_full + 8 * gq);
          if (CLUSTER > 1)
            asm volatile("st.relaxed.cluster.shared::cluster.b32 [%0], %1;" ::"r"(cluster_map(tq_slots + 4 * gq, 1)), "r"(msg)
                         : "memory");
          if ((gq ^= 1) == 0) gph ^= 1;
        } else {
          t = tq.next(false);
        }
        if (t < 0) break;
        TileCoord tc = decode_tile(p, t);
        if (CLUSTER > 1) tc.m_blk = tc.m_blk * CLUSTER + crank;
        for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          const uint32_t sA = smem_base + s * C::STAGE_BYTES;
          const uint32_t sB = sA + C::A_BYTES;
          constexpr int BNL = BN / CLUSTER;           // B rows this CTA loads
          const int n_base = tc.n_blk * BN + crank * BNL;
          if (!PAIR) {
            const uint32_t full = bar_full + 8 * s;
            mbar_expect_tx(full, C::STAGE_BYTES);
            if (!A_MN) {
              tma_load_2d(&tmA, full, sA, kb * BK, tc.m_blk * BM);
            } else {
#pragma unroll
              for (int c = 0; c < BM / 64; ++c)
                tma_load_2d(&tmA, full, sA + c * (BK * 128), tc.m_blk * BM + c * 64, kb * BK);
            }
            if (!B_MN) {
              tma_load_2d(&tmB, full, sB, kb * BK, n_base);
            } else {
#pragma unroll
              for (int c = 0; c < BNL / 64; ++c)
                tma_load_2d(&tmB, full, sB + c * (BK * 128), n_base + c * 64, kb * BK);
            }
          } else {
            // both CTAs' bytes are counted on the leader's barrier (its MMA thread consumes both halves)
            const uint32_t full = cluster_map(bar_full + 8 * s, 0);
            if (crank == 0) mbar_expect_tx(bar_full + 8 * s, 2 * C::STAGE_BYTES);
            if (!A_MN) {
              tma_load_2d_pair(&tmA, full, sA, kb * BK, tc.m_blk * BM);
            } else {
#pragma unroll
              for (int c = 0; c < BM / 64; ++c)
                tma_load_2d_pair(&tmA, full, sA + c * (BK * 128), tc.m_blk * BM + c * 64, kb * BK);
            }
            if (!B_MN) {
              tma_load_2d_pair(&tmB, full, sB, kb * BK, n_base);
            } else {
#pragma unroll
              for (int c = 0; c < BNL / 64; ++c)
                tma_load_2d_pair(&tmB, full, sB + c * (BK * 128), n_base + c * 64, kb * BK);
            }
          }
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
      }
      // every leader draws exactly one item past the end; the last one to do so puts the counters back to zero for the
      // next launch that uses this slot (launches of one stream are ordered; slots are not shared across streams)
      if (crank == 0 && p.dynamic && atomicAdd(p.sched + 1, 1) == n_leaders - 1) {
        p.sched[0] = 0;
        p.sched[1] = 0;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0 && crank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM * CLUSTER, BN, A_MN, B_MN);
      constexpr uint16_t kPairMask = 3;
      // K-major: 8-row groups are 1024 B apart (SBO); LBO unused for swizzled K-major.
      // MN-major: 8-k-row groups are 1024 B apart (SBO); 64-wide MN blocks are BK*128 B apart (LBO).
      constexpr uint32_t A_LBO = A_MN ? BK * 128 : 16, B_LBO = B_MN ? BK * 128 : 16;
      constexpr uint32_t A_KSTEP = A_MN ? 2048 : 32, B_KSTEP = B_MN ? 2048 : 32;
      constexpr uint32_t kDescHi = umma_desc_hi(1024, UMMA_SW128);
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      for (int t = tq.next(false); t >= 0; t = tq.next(false)) {
        const TileCoord tc = decode_tile(p, t);
        mbar_wait(bar_tempty + 8 * as, aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        int sel = (A_MN && p.colsum != nullptr) ? tc.kb0 % p.n_tiles : -1;  // == n_blk: stage goes to the colsum warps
        for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
          mbar_wait(bar_full + 8 * s, ph);
          tc_fence_after();
          const uint32_t sA = smem_base + s * C::STAGE_BYTES;
          const uint32_t sB = sA + C::A_BYTES;
          const uint32_t a_lo = umma_desc_lo(sA, A_LBO), b_lo = umma_desc_lo(sB, B_LBO);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint32_t acc = (kb > tc.kb0 || k > 0) ? 1u : 0u;
            if (PAIR)
              tc_mma_bf16_lh_pair(d_tmem, umma_lo_add(a_lo, k * A_KSTEP), kDescHi, umma_lo_add(b_lo, k * B_KSTEP),
                                  kDescHi, idesc, acc);
            else
              tc_mma_bf16_lh(d_tmem, umma_lo_add(a_lo, k * A_KSTEP), kDescHi, umma_lo_add(b_lo, k * B_KSTEP), kDescHi,
                             idesc, acc);
          }
          // frees the smem stage once these MMAs retire (in both CTAs of a pair); in column-sum mode every
          // n_tiles-th stage goes to the colsum warps first, which release it after reading the A tile
          const uint32_t rel = (sel == tc.n_blk) ? bar_done : bar_empty;
          if (sel >= 0 && ++sel == p.n_tiles) sel = 0;
          if (PAIR) tc_commit_pair(rel + 8 * s, kPairMask);
          else tc_commit(rel + 8 * s);
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
        // accumulator ready for the epilogue (of both CTAs of a pair)
        if (PAIR) tc_commit_pair(bar_tfull + 8 * as, kPairMask);
        else tc_commit(bar_tfull + 8 * as);
        if (++as == 2) {
          as = 0;
          aph ^= 1;
        }
      }
    }
  } else if (EPI == EPI_NONE && warp >= 6) {
    // ===================== column sums of A (warps 6, 7; wgrad only) =====================
    // The n_tiles items that share a row block see the same A tiles: item n_blk sums the k-blocks with
    // kb % n_tiles == n_blk, so the extra shared-memory reads are spread over all items.
    if (cs_active) {
      const int cw = warp - 6;  // 64-column block of the [64 k x 128 m] A tile this warp sums
      // lane reads the 16-byte chunk (lane & 7) of k-rows (lane >> 3) + 4 i; chunks are XOR-swizzled by (row & 7)
      const uint32_t rg = lane >> 3, ch = lane & 7;
      const uint32_t off_even = rg * 128 + ((ch ^ rg) << 4);        // rows with (row & 7) == rg
      const uint32_t off_odd = rg * 128 + ((ch ^ (rg + 4)) << 4);   // rows with (row & 7) == rg + 4
      int s = 0;
      uint32_t done_ph = 0;  // phase bit per stage of the `done` barriers (only some stages use them)
      for (int t = tq.next(true); t >= 0; t = tq.next(true)) {
        TileCoord tc = decode_tile(p, t);
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
        int sel = tc.kb0 % p.n_tiles;
        for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
          if (sel == tc.n_blk) {
            mbar_wait(bar_done + 8 * s, (done_ph >> s) & 1u);
            done_ph ^= 1u << s;
            const uint32_t blk = smem_base + s * C::STAGE_BYTES + cw * (BK * 128);
#pragma unroll
            for (int i = 0; i < BK / 4; ++i) {
              uint32_t u0, u1, u2, u3;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(u0), "=r"(u1), "=r"(u2), "=r"(u3)
                           : "r"(blk + i * 512 + ((i & 1) ? off_odd : off_even)));
              acc[0] += __uint_as_float(u0 << 16);
              acc[1] += __uint_as_float(u0 & 0xffff0000u);
              acc[2] += __uint_as_float(u1 << 16);
              acc[3] += __uint_as_float(u1 & 0xffff0000u);
              acc[4] += __uint_as_float(u2 << 16);
              acc[5] += __uint_as_float(u2 & 0xffff0000u);
              acc[6] += __uint_as_float(u3 << 16);
              acc[7] += __uint_as_float(u3 & 0xffff0000u);
            }
            named_bar_sync(2, 64);  // both colsum warps are done with the stage
            if (warp == 6 && lane == 0) mbar_arrive(bar_empty + 8 * s);
          }
          if (++sel == p.n_tiles) sel = 0;
          if (++s == STAGES) s = 0;
        }
        // fold the four row groups, then lanes 0..7 hold 8 columns each
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 8);
          acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
        }
        if (lane < 8) {
          const int m_blk = CLUSTER > 1 ? tc.m_blk * CLUSTER + crank : tc.m_blk;
          const int m = m_blk * BM + cw * 64 + 8 * lane;
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (m + e < p.M) atomicAdd(p.colsum + m + e, acc[e] * p.alpha);
        }
      }
    }
  } else if (EPI != EPI_NONE) {
    // ===================== fused epilogue (warps 2..9 = 256 threads) =====================
    // The activation costs ~13 (forward) / ~20 (backward) instructions per element, and a lone warp per scheduler
    // issues a dependent stream at ~0.25 instructions per clock (measured: four epilogue warps needed 23 k cycles for
    // a 128 x 256 tile against the 5.8 k of a K = 768 mainloop).  Eight warps: two per TMEM lane quadrant, in two
    // groups g = (warp - 2) / 4.  EPI_MUL: group g takes the 64-column chunks of parity g (the multiplier chunk sits in
    // its own staging block).  EPI_GELU_FWD: both groups share a chunk (32 columns each) and chunks alternate between
    // two staging sets, see below.
    const int ew = warp - 2;       // 0..7
    const int grp = ew >> 2;       // chunk parity
    const int quad = warp & 3;     // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;
    const int etid = ew * 32 + lane;
    const bool store_thread = (ew & 3) == 0 && lane == 0;
    int as = 0;
    uint32_t aph = 0, gph = 0;
    uint32_t ek0 = p.epi.k0, ek1 = p.epi.k1;
    if (p.epi.thresh != 0 && p.epi.epoch != nullptr) {
      const unsigned long long e = *p.epi.epoch;
      ek0 ^= epi_fmix32((uint32_t)e * 0x9E3779B1u + 0x68E31DA4u);
      ek1 += epi_fmix32((uint32_t)(e >> 32) ^ 0xB5297A4Du) + (uint32_t)e;
    }
    constexpr int NCHUNK = BN / 64;
    constexpr int UNITS = NCHUNK;  // 32-column units this group handles per tile: NCHUNK / 2 chunks x 2 halves
    for (int t = tq.next(true); t >= 0; t = tq.next(true)) {
      TileCoord tc = decode_tile(p, t);
      if (CLUSTER > 1) tc.m_blk = tc.m_blk * CLUSTER + crank;
      const int n0 = tc.n_blk * BN;
      const int m_glob = tc.m_blk * BM + row;  // global output row of this thread
      // EPI_MUL: the multiplier tile G[128 rows x 256 cols] arrives by TMA, one [128 x 64] chunk per staging block
      // (this group's two chunks -> blocks grp and 2 + grp), requested before the accumulator is waited for.  (Per-
      // thread global loads of a row segment are 32-way divergent: 4096 LSU cycles per tile, measured +38 us per
      // launch; the bulk copy costs nothing but its bytes.)  The product is written back in place and stored.
      if (EPI == EPI_MUL && store_thread) {
        tma_wait_group_read0();  // the previous tile's stores have finished reading both blocks
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int cbk = grp + 2 * k;
          if (n0 + cbk * 64 < p.N) {
            mbar_expect_tx(bar_g + 8 * cbk, BM * 128);
            tma_load_2d(&tmD2, bar_g + 8 * cbk, sC + cbk * (BM * 128), n0 + cbk * 64, tc.m_blk * BM);
          } else {
            mbar_arrive(bar_g + 8 * cbk);
          }
        }
      }
      mbar_wait(bar_tfull + 8 * as, aph);
      tc_fence_after();
      float* bs_t = bias_s + as * BN;  // per accumulator stage: the other group may still be reading the previous tile's
      for (int i = etid; i < BN; i += kEpiWarps * 32) {
        const int n = n0 + i;
        bs_t[i] = (p.bias != nullptr && n < p.N) ? p.bias[n] : 0.f;
      }
      named_bar_sync(1, kEpiWarps * 32);
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN;
      if (EPI == EPI_GELU_FWD) {
        // All four warps of a TMEM lane quadrant work on the SAME 64-column chunk (warp `sub` = ew / 4 takes columns
        // [16 sub, 16 sub + 16) of its row); consecutive chunks alternate between two staging sets (H block c & 1, G
        // block 2 + (c & 1)).  The bulk stores of chunk c - 1 get the whole arithmetic of chunk c to finish reading
        // their set before chunk c + 1 overwrites it, so the wait in front of the barrier is free.
        const bool st_thread = ew == 0 && lane == 0;
        const int sub = ew >> 2;  // 0..3
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) {
          const uint32_t row_h = sC + (c & 1) * (BM * 128) + row * 128;
          const uint32_t row_g = row_h + 2 * (BM * 128);
          uint32_t r[16];
          tmem_ld16(t_addr + c * 64 + sub * 16, r);
          tmem_ld_wait();
          if (c == NCHUNK - 1) {  // this warp's last read of the accumulator stage: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (PAIR) mbar_arrive_cluster_relaxed(cluster_map(bar_tempty + 8 * as, 0));
              else mbar_arrive(bar_tempty + 8 * as);
            }
          }
          // keep bits of this thread's 16 columns: half of the word of the 32-column block they lie in
          const uint32_t kw = epi_keep_word(p.epi, ek0, ek1, (uint32_t)m_glob, (uint32_t)((n0 + c * 64) >> 5) + (sub >> 1)) >>
                              (16 * (sub & 1));
          const float* bs = bs_t + c * 64 + sub * 16;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const uint32_t sw = ((((sub * 2 + q) ^ (row & 7)) & 7) << 4);
            float hv[8], gv[8];
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
              const float2 v = __ffma2_rn(make_float2(__uint_as_float(r[8 * q + e]), __uint_as_float(r[8 * q + e + 1])),
                                          make_float2(p.alpha, p.alpha), make_float2(bs[8 * q + e], bs[8 * q + e + 1]));
              const float2 dm = make_float2((kw & (1u << (8 * q + e))) ? p.epi.inv_keep : 0.f,
                                            (kw & (2u << (8 * q + e))) ? p.epi.inv_keep : 0.f);
              float2 val = v, grad = v;
              gelu_pair(v, val, grad);
              const float2 h2 = __fmul2_rn(val, dm), g2 = __fmul2_rn(grad, dm);
              hv[e] = h2.x;
              hv[e + 1] = h2.y;
              gv[e] = g2.x;
              gv[e + 1] = g2.y;
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_h + sw), "r"(pack_bf16(hv[0], hv[1])),
                         "r"(pack_bf16(hv[2], hv[3])), "r"(pack_bf16(hv[4], hv[5])), "r"(pack_bf16(hv[6], hv[7]))
                         : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_g + sw), "r"(pack_bf16(gv[0], gv[1])),
                         "r"(pack_bf16(gv[2], gv[3])), "r"(pack_bf16(gv[4], gv[5])), "r"(pack_bf16(gv[6], gv[7]))
                         : "memory");
          }
          fence_proxy_async_smem();
          if (st_thread) tma_wait_group_read0();  // chunk c - 1's stores (the OTHER set) have finished reading
          named_bar_sync(2, kEpiWarps * 32);
          if (st_thread) {
            if (n0 + c * 64 < p.N) {
              tma_store_2d(&tmD, sC + (c & 1) * (BM * 128), n0 + c * 64, tc.m_blk * BM);
              tma_store_2d(&tmD2, sC + (2 + (c & 1)) * (BM * 128), n0 + c * 64, tc.m_blk * BM);
            }
            tma_commit_group();
          }
        }
      } else {
#pragma unroll
        for (int u = 0; u < UNITS; ++u) {
          const int cb = grp + 2 * (u >> 1), hf = u & 1;
          const uint32_t blk_u = sC + cb * (BM * 128) + row * 128;
          if (hf == 0) mbar_wait(bar_g + 8 * cb, gph);  // multiplier chunk landed (a thread touches only its own row of it)
          uint32_t r[32];
          tmem_ld32(t_addr + cb * 64 + hf * 32, r);
          tmem_ld_wait();
          if (u == UNITS - 1) {  // this warp's last read of the accumulator stage: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (PAIR) mbar_arrive_cluster_relaxed(cluster_map(bar_tempty + 8 * as, 0));
              else mbar_arrive(bar_tempty + 8 * as);
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t sw = ((((hf * 4 + q) ^ (row & 7)) & 7) << 4);
            uint32_t zw[4];
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(zw[0]), "=r"(zw[1]), "=r"(zw[2]), "=r"(zw[3])
                         : "r"(blk_u + sw));
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 gm = make_float2(__uint_as_float(zw[e] << 16), __uint_as_float(zw[e] & 0xffff0000u));
              const float2 d = __fmul2_rn(__fmul2_rn(make_float2(__uint_as_float(r[8 * q + 2 * e]),
                                                                 __uint_as_float(r[8 * q + 2 * e + 1])),
                                                     make_float2(p.alpha, p.alpha)), gm);
              o[e] = pack_bf16(d.x, d.y);
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(blk_u + sw), "r"(o[0]), "r"(o[1]), "r"(o[2]),
                         "r"(o[3])
                         : "memory");
          }
          if (hf == 1) {
            fence_proxy_async_smem();
            named_bar_sync(2 + grp, 128);
            if (store_thread) {
              if (n0 + cb * 64 < p.N) tma_store_2d(&tmD, sC + cb * (BM * 128), n0 + cb * 64, tc.m_blk * BM);
              tma_commit_group();
            }
          }
        }
      }
      gph ^= 1;
      if (++as == 2) {
        as = 0;
        aph ^= 1;
      }
    }
    if (store_thread) tma_wait_group0();
  } else {
    // ===================== epilogue (warps 2..5 = 128 threads) =====================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;
    const int etid = threadIdx.x - 64;  // 0..127
    const bool store_thread = (etid == 0);
    int as = 0;
    uint32_t aph = 0;
    for (int t = tq.next(true); t >= 0; t = tq.next(true)) {
      TileCoord tc = decode_tile(p, t);
      if (CLUSTER > 1) tc.m_blk = tc.m_blk * CLUSTER + crank;
      const int n0 = tc.n_blk * BN;
      mbar_wait(bar_tfull + 8 * as, aph);
      tc_fence_after();
      for (int i = etid; i < BN; i += 128) {
        const int n = n0 + i;
        bias_s[i] = (p.bias != nullptr && n < p.N && tc.kb0 == 0) ? p.bias[n] : 0.f;
      }
      named_bar_sync(1, 128);
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN;
      constexpr int CW = OUT_F32 ? 32 : 64;    // output columns per 128-byte staging block
      constexpr int NCHUNK = BN / CW;
#pragma unroll 1
      for (int cb = 0; cb < NCHUNK; ++cb) {
        const uint32_t blk = sC + (cb & 1) * (BM * 128) + row * 128;
        // the TMA store that last read this staging block (two chunks ago) must have finished reading it
        if (store_thread) tma_wait_group_read1();
        named_bar_sync(1, 128);
        if (OUT_F32) {
          uint32_t r[32];
          tmem_ld32(t_addr + cb * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const uint32_t dst = blk + (((q ^ (row & 7)) & 7) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst),
                         "f"(__uint_as_float(r[4 * q]) * p.alpha + bias_s[cb * 32 + 4 * q]),
                         "f"(__uint_as_float(r[4 * q + 1]) * p.alpha + bias_s[cb * 32 + 4 * q + 1]),
                         "f"(__uint_as_float(r[4 * q + 2]) * p.alpha + bias_s[cb * 32 + 4 * q + 2]),
                         "f"(__uint_as_float(r[4 * q + 3]) * p.alpha + bias_s[cb * 32 + 4 * q + 3])
                         : "memory");
          }
        } else {
          uint32_t r0[32], r1[32];
          tmem_ld32(t_addr + cb * 64, r0);
          tmem_ld32(t_addr + cb * 64 + 32, r1);
          tmem_ld_wait();
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const uint32_t(&r)[32] = hf == 0 ? r0 : r1;
            const float* bs = bias_s + cb * 64 + hf * 32;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int chunk = hf * 4 + q;
              const uint32_t dst = blk + (((chunk ^ (row & 7)) & 7) << 4);
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[8 * q + e]) * p.alpha + bs[8 * q + e];
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pack_bf16(v[0], v[1])),
                           "r"(pack_bf16(v[2], v[3])), "r"(pack_bf16(v[4], v[5])), "r"(pack_bf16(v[6], v[7]))
                           : "memory");
            }
          }
        }
        if (cb == NCHUNK - 1) {  // TMEM stage drained: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR) mbar_arrive_cluster_relaxed(cluster_map(bar_tempty + 8 * as, 0));
            else mbar_arrive(bar_tempty + 8 * as);
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (store_thread) {
          if (n0 + cb * CW < p.N) {
            if (OUT_F32)
              tma_reduce_add_2d(&tmD, sC + (cb & 1) * (BM * 128), n0 + cb * CW, tc.m_blk * BM);
            else
              tma_store_2d(&tmD, sC + (cb & 1) * (BM * 128), n0 + cb * CW, tc.m_blk * BM);
          }
          tma_commit_group();  // (an empty group when the chunk lies past N keeps the wait_group arithmetic uniform)
        }
      }
      if (++as == 2) {
        as = 0;
        aph ^= 1;
      }
    }
    if (store_thread) tma_wait_group0();
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // no CTA leaves while its peer may still signal it or read its B half
  if (warp == 1) {
    if (PAIR) tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
    else tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// -----------------------------------------------------------------------------------------------
// host launcher
// -----------------------------------------------------------------------------------------------
// Work-queue counters: a ring of (next item, drained leaders) pairs in device global memory, zero between launches
// (the kernel's last leader resets its pair).  A launch takes the next pair of the ring; 4096 pairs are far more than
// the launches that can be in flight at once, and a captured graph keeps the pairs it was captured with.
constexpr int kSchedSlots = 4096;
__device__ int g_sched[2 * kSchedSlots];

int* next_sched_slot() {
  static std::atomic<unsigned> seq{0};
  static std::mutex mu;
  static int* base[64] = {nullptr};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  int* b = nullptr;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (base[dev] == nullptr) {
      void* ptr = nullptr;
      if (cudaGetSymbolAddress(&ptr, g_sched) != cudaSuccess) return nullptr;
      base[dev] = static_cast<int*>(ptr);
    }
    b = base[dev];
  }
  return b + 2 * (seq.fetch_add(1) % kSchedSlots);
}

struct EpiArgs {  // host-side description of a fused epilogue (EPI_NONE: all zero)
  void* d2 = nullptr;       // EPI_GELU_FWD: second output H [M, ldd2]
  int64_t ldd2 = 0;
  const void* z = nullptr;  // EPI_MUL: elementwise multiplier G [M, ldz]
  int64_t ldz = 0;
  float p_drop = 0.f;
  uint64_t seed = 0, offset = 0;
  const uint64_t* epoch = nullptr;
};

void fill_epi(EpiParams& e, const EpiArgs& a) {
  e.z = (const __nv_bfloat16*)a.z;
  e.ldz = a.ldz;
  uint64_t zz = a.seed * 0x9E3779B97F4A7C15ull + a.offset * 0xD1B54A32D192ED03ull + 0x2545F4914F6CDD1Dull;
  zz = (zz ^ (zz >> 30)) * 0xBF58476D1CE4E5B9ull;
  zz = (zz ^ (zz >> 27)) * 0x94D049BB133111EBull;
  zz ^= zz >> 31;
  e.k0 = (uint32_t)zz;
  e.k1 = (uint32_t)(zz >> 32);
  e.epoch = reinterpret_cast<const unsigned long long*>(a.epoch);
  const double full = (double)(1u << kEpiDropBits);
  double t = (double)a.p_drop * full + 0.5;
  if (t < 1.0) t = 1.0;
  e.thresh = a.p_drop > 0.f ? (uint32_t)(t > full - 1.0 ? full - 1.0 : t) : 0u;
  for (int i = 0; i < kEpiDropBits; ++i) e.tmask[i] = ((e.thresh >> i) & 1u) ? 0xFFFFFFFFu : 0u;
  e.inv_keep = a.p_drop > 0.f ? (float)(full / (full - (double)e.thresh)) : 1.0f;
}

template <int BN, bool A_MN, bool B_MN, bool OUT_F32, int CLUSTER, int EPI = EPI_NONE>
int launch_impl(const void* A, int64_t lda, const void* B, int64_t ldb, void* D, int64_t ldd,
                const float* bias, float* colsum, float alpha, int64_t M, int64_t N, int64_t K, int k_splits_req,
                cudaStream_t stream, const EpiArgs& ea = EpiArgs()) {
  using C = Cfg<BN, OUT_F32, CLUSTER, EPI>;
  auto kern = gemm_kernel<BN, A_MN, B_MN, OUT_F32, CLUSTER, EPI>;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kern), C::SMEM_BYTES)) return rc;
  CUtensorMap tmA, tmB, tmD;
  int rc;
  if (!A_MN)
    rc = make_tmap_2d(&tmA, A, 2, false, K, M, lda * 2, BK, BM, SWZ_128);
  else
    rc = make_tmap_2d(&tmA, A, 2, false, M, K, lda * 2, 64, BK, SWZ_128);
  if (rc) return rc;
  if (!B_MN)
    rc = make_tmap_2d(&tmB, B, 2, false, K, N, ldb * 2, BK, BN / CLUSTER, SWZ_128);  // per-CTA share of the tile
  else
    rc = make_tmap_2d(&tmB, B, 2, false, N, K, ldb * 2, 64, BK, SWZ_128);
  if (rc) return rc;
  if (OUT_F32)
    rc = make_tmap_2d(&tmD, D, 4, true, N, M, ldd * 4, 32, BM, SWZ_128);
  else
    rc = make_tmap_2d(&tmD, D, 2, false, N, M, ldd * 2, 64, BM, SWZ_128);
  if (rc) return rc;
  CUtensorMap tmD2 = tmD;
  if (EPI == EPI_GELU_FWD) {
    rc = make_tmap_2d(&tmD2, ea.d2, 2, false, N, M, ea.ldd2 * 2, 64, BM, SWZ_128);
    if (rc) return rc;
  } else if (EPI == EPI_MUL) {  // the multiplier G is read through the same [128 x 64] boxes the product leaves by
    rc = make_tmap_2d(&tmD2, ea.z, 2, false, N, M, ea.ldz * 2, 64, BM, SWZ_128);
    if (rc) return rc;
  }

  GemmParams p;
  fill_epi(p.epi, ea);
  // Measured (2 x B200, cfg3 step): with the static stride the data-parallel step is 46.5 ms against 46.3 ms on one
  // GPU; drawing items from the counter makes it 48.2 ms (the GPU-scope atomics queue behind the NVLink traffic of the
  // overlapped all-reduce).  What used to cost 1.8 ms per step under data parallelism was not SM contention but the
  // MEMBAR.ALL.GPU of the `.release.cluster` hand-offs, which has to drain behind that same traffic.  Static is the
  // default; SCT_GEMM_DYNAMIC=1 selects the counter.
  p.dynamic = env_int("SCT_GEMM_DYNAMIC", 0);
  p.sched = next_sched_slot();
  SCT_CHECK(p.sched != nullptr, "work-queue counters unavailable");
  p.M = (int)M;
  p.N = (int)N;
  p.K = (int)K;
  p.m_tiles = (int)((M + BM * CLUSTER - 1) / (BM * CLUSTER));  // CLUSTER = 2: pairs of vertically adjacent tiles
  p.n_tiles = (int)((N + BN - 1) / BN);
  p.kb_total = (int)((K + BK - 1) / BK);
  const int sms = num_sms() / CLUSTER;  // concurrent work items (CTAs, or CTA pairs)
  int ks = 1;
  if (OUT_F32) {
    // split-K so that a small [N_out x K_in] weight gradient still fills the machine: the smallest split count
    // whose work items fill >= 92 % of the last wave of 148 CTAs (108 tiles x 2 splits = 1.46 waves ran at 779
    // TFLOP/s, x 4 = 2.9 waves at 917), keeping >= 8 k-blocks per item; every extra split adds reduce-add traffic.
    const int tiles = p.m_tiles * p.n_tiles;
    if (k_splits_req > 0) {
      ks = k_splits_req;
    } else {
      int best = 1;
      double best_eff = 0.0;
      const int ks_max = p.kb_total / 8 > 0 ? p.kb_total / 8 : 1;
      for (int c = 1; c <= 32 && c <= ks_max; ++c) {
        const int total = tiles * c;
        const int waves = (total + sms - 1) / sms;
        const double eff = (double)total / ((double)waves * sms);
        if (eff > best_eff + 1e-9) {
          best_eff = eff;
          best = c;
        }
        if (eff >= 0.92) break;
      }
      ks = best;
    }
    if (ks < 1) ks = 1;
    if (ks > p.kb_total) ks = p.kb_total;
  }
  p.kb_per_split = (p.kb_total + ks - 1) / ks;
  p.k_splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.n_fast = (p.n_tiles <= p.m_tiles) ? 1 : 0;
  p.bias = bias;
  p.colsum = colsum;
  p.alpha = alpha;
  const int total = p.m_tiles * p.n_tiles * p.k_splits;
  if (CLUSTER == 1) {
    const int grid = total < sms ? total : sms;
    kern<<<grid, C::THREADS, C::SMEM_BYTES, stream>>>(tmA, tmB, tmD, tmD2, p);
  } else {
    const int clusters = total < sms ? total : sms;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * CLUSTER));
    cfg.blockDim = dim3(C::THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CLUSTER;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SCT_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmD, tmD2, p));
  }
  SCT_LAUNCH_CHECK();
  return 0;
}

// CTA pairs (cta_group::2) for the 256-wide tile whenever there are at least two row tiles; SCT_GEMM_PAIR=0 forces
// the single-CTA kernel (A/B timing).
template <int BN, bool A_MN, bool B_MN, bool OUT_F32>
int launch(const void* A, int64_t lda, const void* B, int64_t ldb, void* D, int64_t ldd, const float* bias,
           float* colsum, float alpha, int64_t M, int64_t N, int64_t K, int k_splits_req, cudaStream_t stream) {
  const int use_cluster = env_int("SCT_GEMM_PAIR", 1);
  if constexpr (BN == 256) {
    if (use_cluster && M > BM)
      return launch_impl<BN, A_MN, B_MN, OUT_F32, 2>(A, lda, B, ldb, D, ldd, bias, colsum, alpha, M, N, K,
                                                     k_splits_req, stream);
  }
  return launch_impl<BN, A_MN, B_MN, OUT_F32, 1>(A, lda, B, ldb, D, ldd, bias, colsum, alpha, M, N, K, k_splits_req,
                                                 stream);
}

template <bool B_MN, int EPI>
int launch_epi(const void* A, int64_t lda, const void* B, int64_t ldb, void* D, int64_t ldd, const float* bias,
               float alpha, int64_t M, int64_t N, int64_t K, cudaStream_t stream, const EpiArgs& ea) {
  if (env_int("SCT_GEMM_PAIR", 1) && M > BM)
    return launch_impl<256, false, B_MN, false, 2, EPI>(A, lda, B, ldb, D, ldd, bias, nullptr, alpha, M, N, K, 1, stream, ea);
  return launch_impl<256, false, B_MN, false, 1, EPI>(A, lda, B, ldb, D, ldd, bias, nullptr, alpha, M, N, K, 1, stream, ea);
}

int check_common(const void* A, const void* B, const void* D, int64_t M, int64_t N, int64_t K) {
  SCT_CHECK(A && B && D, "null operand pointer");
  SCT_CHECK(M > 0 && N > 0 && K > 0, "empty GEMM (M=%lld N=%lld K=%lld)", (long long)M, (long long)N,
            (long long)K);
  SCT_CHECK(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "dimension overflow");
  return 0;
}

__global__ void read_flag_kernel(int* out) { *out = g_timeout_flag; }

}  // namespace

int gemm_timeout_flag() {
  int* d = nullptr;
  int h = 0;
  if (cudaMalloc(&d, sizeof(int)) != cudaSuccess) return -1;
  read_flag_kernel<<<1, 1>>>(d);
  cudaMemcpy(&h, d, sizeof(int), cudaMemcpyDeviceToHost);
  cudaFree(d);
  return h;
}

}  // namespace sct

extern "C" {

int32_t sct_gemm_bf16_nt(const void* A, int64_t lda, const void* W, int64_t ldw, void* D, int64_t ldd,
                         const float* bias, float alpha, int64_t M, int64_t N, int64_t K, int32_t bn,
                         void* stream) {
  if (int rc = sct::check_common(A, W, D, M, N, K)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (bn == 256) return sct::launch<256, false, false, false>(A, lda, W, ldw, D, ldd, bias, nullptr, alpha, M, N, K, 1, st);
  // 128x64 tiles: skinny GEMMs of the decode step (M <= 128 rows): N / 64 CTAs stream the weight instead of N / 256
  if (bn == 64) return sct::launch<64, false, false, false>(A, lda, W, ldw, D, ldd, bias, nullptr, alpha, M, N, K, 1, st);
  return sct::launch<128, false, false, false>(A, lda, W, ldw, D, ldd, bias, nullptr, alpha, M, N, K, 1, st);
}

int32_t sct_gemm_bf16_nn(const void* A, int64_t lda, const void* W, int64_t ldw, void* D, int64_t ldd,
                         const float* bias, float alpha, int64_t M, int64_t N, int64_t K, int32_t bn,
                         void* stream) {
  if (int rc = sct::check_common(A, W, D, M, N, K)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (bn == 256) return sct::launch<256, false, true, false>(A, lda, W, ldw, D, ldd, bias, nullptr, alpha, M, N, K, 1, st);
  return sct::launch<128, false, true, false>(A, lda, W, ldw, D, ldd, bias, nullptr, alpha, M, N, K, 1, st);
}

int32_t sct_gemm_bf16_nt_gelu(const void* A, int64_t lda, const void* W, int64_t ldw, void* H, int64_t ldh, void* G,
                              int64_t ldg, const float* bias, int64_t M, int64_t N, int64_t K, float p_drop,
                              uint64_t seed, uint64_t offset, const uint64_t* epoch, void* stream) {
  if (int rc = sct::check_common(A, W, H, M, N, K)) return rc;
  SCT_CHECK(G != nullptr, "null G output");
  SCT_CHECK(N % 64 == 0, "fused GELU epilogue needs N %% 64 == 0 (N = %lld)", (long long)N);
  SCT_CHECK(p_drop >= 0.f && p_drop < 1.f, "p_drop out of range");
  sct::EpiArgs ea;
  ea.d2 = G; ea.ldd2 = ldg; ea.p_drop = p_drop; ea.seed = seed; ea.offset = offset; ea.epoch = epoch;
  return sct::launch_epi<false, sct::EPI_GELU_FWD>(A, lda, W, ldw, H, ldh, bias, 1.0f, M, N, K,
                                                   static_cast<cudaStream_t>(stream), ea);
}

int32_t sct_gemm_bf16_nn_mul(const void* A, int64_t lda, const void* W, int64_t ldw, const void* G, int64_t ldg,
                             void* D, int64_t ldd, int64_t M, int64_t N, int64_t K, void* stream) {
  if (int rc = sct::check_common(A, W, D, M, N, K)) return rc;
  SCT_CHECK(G != nullptr, "null G input");
  SCT_CHECK(N % 64 == 0 && ldg % 8 == 0, "fused multiply epilogue needs N %% 64 == 0 and ldg %% 8 == 0");
  SCT_CHECK((reinterpret_cast<uintptr_t>(G) & 15) == 0, "G must be 16-byte aligned");
  sct::EpiArgs ea;
  ea.z = G; ea.ldz = ldg;
  return sct::launch_epi<true, sct::EPI_MUL>(A, lda, W, ldw, D, ldd, nullptr, 1.0f, M, N, K,
                                             static_cast<cudaStream_t>(stream), ea);
}

int32_t sct_gemm_bf16_tn(const void* A, int64_t lda, const void* B, int64_t ldb, float* D, int64_t ldd,
                         float alpha, int64_t M, int64_t N, int64_t K, int32_t k_splits, void* stream) {
  return sct_gemm_bf16_tn_colsum(A, lda, B, ldb, D, ldd, nullptr, alpha, M, N, K, k_splits, stream);
}

int32_t sct_gemm_bf16_tn_colsum(const void* A, int64_t lda, const void* B, int64_t ldb, float* D, int64_t ldd,
                                float* colsum, float alpha, int64_t M, int64_t N, int64_t K, int32_t k_splits,
                                void* stream) {
  if (int rc = sct::check_common(A, B, D, M, N, K)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int tn_bn = sct::env_int("SCT_GEMM_TN_BN", 256);  // =128 forces the narrow tile (A/B timing)
  if (tn_bn == 256 && N >= 512)
    return sct::launch<256, true, true, true>(A, lda, B, ldb, D, ldd, nullptr, colsum, alpha, M, N, K, k_splits, st);
  return sct::launch<128, true, true, true>(A, lda, B, ldb, D, ldd, nullptr, colsum, alpha, M, N, K, k_splits, st);
}

}  // extern "C"
