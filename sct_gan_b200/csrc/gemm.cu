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

#include "../../include/sct_b200.h"
#include "common.cuh"

namespace sct {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kThreads = 256;
constexpr int kSmemLimit = 232448;  // 227 KB

struct GemmParams {
  int M, N, K;
  int m_tiles, n_tiles, k_splits;
  int kb_total, kb_per_split;
  int n_fast;         // 1: consecutive tiles walk N first (A tile reused from L2), 0: walk M first
  const float* bias;  // nullable, fp32 [N]
  float alpha;        // output scale applied before bias
  float* colsum;      // nullable (MN-major A only), fp32 [M]: += alpha * sum_k A[k, m]
};

template <int BN, bool OUT_F32, int CLUSTER>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN / CLUSTER * BK * 2;  // this CTA's share of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // Epilogue staging: two [128 rows x 128 B] blocks (64 bf16 / 32 fp32 columns each), ping-ponged chunk by chunk.
  // A full-tile staging buffer (64 KB at BN = 256) would leave only 3 pipeline stages, and the mainloop is
  // bound by the bytes it can keep in flight (measured: the MMA thread waited on `full` 44 % of the time).
  static constexpr int C_BYTES = 2 * BM * 128;
  static constexpr int AUX_BYTES = 1024;  // barriers + tmem ptr
  static constexpr int STAGES_RAW = (kSmemLimit - 1024 - C_BYTES - AUX_BYTES - BN * 4) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + C_BYTES + AUX_BYTES + BN * 4;
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

template <int BN, bool A_MN, bool B_MN, bool OUT_F32, int CLUSTER>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmD, const GemmParams p) {
  using C = Cfg<BN, OUT_F32, CLUSTER>;
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
  volatile uint32_t* tmem_ptr_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_addr - smem_base));
  float* bias_s = reinterpret_cast<float*>(smem_gen + (sAux + C::AUX_BYTES - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // work items: with CLUSTER = 2 an item is a PAIR of vertically adjacent tiles (p.m_tiles then counts pairs)
  const int total_tiles = p.m_tiles * p.n_tiles * p.k_splits;
  const int crank = CLUSTER > 1 ? (int)cluster_ctarank() : 0;
  const int item0 = CLUSTER > 1 ? (int)(blockIdx.x / CLUSTER) : (int)blockIdx.x;
  const int item_stride = CLUSTER > 1 ? (int)(gridDim.x / CLUSTER) : (int)gridDim.x;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmD);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
      mbar_init(bar_done + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tempty + 8 * s, 4 * CLUSTER);  // one arrival per epilogue warp (of both CTAs of a pair)
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

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = item0; t < total_tiles; t += item_stride) {
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
      for (int t = item0; t < total_tiles; t += item_stride) {
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
  } else if (warp >= 6) {
    // ===================== column sums of A (warps 6, 7; wgrad only) =====================
    // The n_tiles items that share a row block see the same A tiles: item n_blk sums the k-blocks with
    // kb % n_tiles == n_blk, so the extra shared-memory reads are spread over all items.
    if (A_MN && p.colsum != nullptr) {
      const int cw = warp - 6;  // 64-column block of the [64 k x 128 m] A tile this warp sums
      // lane reads the 16-byte chunk (lane & 7) of k-rows (lane >> 3) + 4 i; chunks are XOR-swizzled by (row & 7)
      const uint32_t rg = lane >> 3, ch = lane & 7;
      const uint32_t off_even = rg * 128 + ((ch ^ rg) << 4);        // rows with (row & 7) == rg
      const uint32_t off_odd = rg * 128 + ((ch ^ (rg + 4)) << 4);   // rows with (row & 7) == rg + 4
      int s = 0;
      uint32_t done_ph = 0;  // phase bit per stage of the `done` barriers (only some stages use them)
      for (int t = item0; t < total_tiles; t += item_stride) {
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
  } else {
    // ===================== epilogue (warps 2..5 = 128 threads) =====================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;
    const int etid = threadIdx.x - 64;  // 0..127
    const bool store_thread = (etid == 0);
    int as = 0;
    uint32_t aph = 0;
    for (int t = item0; t < total_tiles; t += item_stride) {
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
            if (PAIR) mbar_arrive_cluster(cluster_map(bar_tempty + 8 * as, 0));
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
template <int BN, bool A_MN, bool B_MN, bool OUT_F32, int CLUSTER>
int launch_impl(const void* A, int64_t lda, const void* B, int64_t ldb, void* D, int64_t ldd,
                const float* bias, float* colsum, float alpha, int64_t M, int64_t N, int64_t K, int k_splits_req,
                cudaStream_t stream) {
  using C = Cfg<BN, OUT_F32, CLUSTER>;
  auto kern = gemm_kernel<BN, A_MN, B_MN, OUT_F32, CLUSTER>;
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

  GemmParams p;
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
    kern<<<grid, kThreads, C::SMEM_BYTES, stream>>>(tmA, tmB, tmD, p);
  } else {
    const int clusters = total < sms ? total : sms;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * CLUSTER));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CLUSTER;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SCT_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmD, p));
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
