// bf16 GEMM for sm_100a: TMA -> shared (128B swizzle) -> tcgen05.mma -> TMEM -> fused epilogue.
//
//   C[M,N] (+)= act( A[M,K] * B[N,K]^T + bias[N] ) + residual[M,N]
//
// A and B can each be K-major (row-major [rows][K]) or MN-major (stored [K][rows]), so the
// same kernel serves linear / 1x1-conv forward (A=x, B=W), dgrad (A=dy, B=W as MN-major) and
// wgrad (A=dy^T, B=x^T, both MN-major, split-K with fp32 atomics) without any transpose pass.
//
// Structure (persistent, warp-specialised, one CTA per SM):
//   warp 0      : TMA producer   (cp.async.bulk.tensor.2d, mbarrier complete_tx)
//   warp 1      : TMEM allocator + MMA issuer (one elected lane issues tcgen05.mma, UMMA 128xBNx16,
//                 accumulator in TMEM, tcgen05.commit releases smem stages / publishes the tile)
//   warps 2..9  : epilogue (tcgen05.ld 32x32b.x32 -> registers -> bias/act/residual -> swizzled smem ->
//                 TMA bulk store); two warps per TMEM lane quarter, each taking half of the columns
//   smem ring of kStages {A 128x64, B BNx64} tiles; TMEM double-buffered accumulators so the
//   epilogue of tile i overlaps the MMAs of tile i+1.
//
// This replaces the cuBLAS addmm / cuDNN 1x1-conv calls on the model zoo's hot path
// (SURVEY.md §2.2 N13, §2.6 S5/S7, §7.1 step 9).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "sm100_common.cuh"

namespace {


template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_z,
                 const GemmParams p) {
  using C = Cfg<BN>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::STAGES * C::A_BYTES;
  uint8_t* smem_store = smem + C::STAGES * C::STAGE_BYTES;   // 1024B-aligned staging for TMA stores
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_store + C::STORE_BYTES);
  uint64_t* full_bar = bars;                     // [STAGES]
  uint64_t* empty_bar = bars + C::STAGES;        // [STAGES]
  uint64_t* tmem_full = bars + 2 * C::STAGES;    // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_stats = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);
  const bool want_stats = p.stats != nullptr && p.out_mode == 0 && p.tma_store;
  if (want_stats) stats_zero(s_stats, NUM_THREADS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (p.tma_store) tma_prefetch_desc(&map_c);
    if (p.tma_store && p.preact != nullptr) tma_prefetch_desc(&map_z);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], EPI_WARPS);   // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_base_slot)),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  const int tiles = p.num_m_blocks * p.num_n_blocks;
  const int work_items = tiles * p.splits;
  const int kb_per_split = (p.num_k_blocks + p.splits - 1) / p.splits;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < work_items; w += gridDim.x) {
        const int tile = w % tiles, split = w / tiles;
        const int m_idx = (p.n_fastest ? tile / p.num_n_blocks : tile % p.num_m_blocks) * BLOCK_M;
        const int n_idx = (p.n_fastest ? tile % p.num_n_blocks : tile / p.num_m_blocks) * BN;
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb0 + kb_per_split, p.num_k_blocks);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          uint8_t* sa = smem_a + stage * C::A_BYTES;
          uint8_t* sb = smem_b + stage * C::B_BYTES;
          const int k_idx = kb * BLOCK_K;
          if (!A_MN) {
            tma_load_2d(&map_a, &full_bar[stage], sa, k_idx, m_idx);           // box {64 k, 128 m}
          } else {
#pragma unroll
            for (int c = 0; c < BLOCK_M / 64; ++c)                                // box {64 m, 64 k}
              tma_load_2d(&map_a, &full_bar[stage], sa + c * (BLOCK_K * 128), m_idx + 64 * c, k_idx);
          }
          if (!B_MN) {
            tma_load_2d(&map_b, &full_bar[stage], sb, k_idx, n_idx);           // box {64 k, BN n}
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_2d(&map_b, &full_bar[stage], sb + c * (BLOCK_K * 128), n_idx + 64 * c, k_idx);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer (one elected thread) ============================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc<BN, A_MN, B_MN>();
      constexpr uint32_t KSTEP_A = A_MN ? ((UMMA_K * 128) >> 4) : ((UMMA_K * 2) >> 4);
      constexpr uint32_t KSTEP_B = B_MN ? ((UMMA_K * 128) >> 4) : ((UMMA_K * 2) >> 4);
      const uint64_t a0 = (A_MN ? make_desc_base(BLOCK_K * 128, 1024) : make_desc_base(16, 1024)) +
                          desc_addr(smem_u32(smem_a));
      const uint64_t b0 = (B_MN ? make_desc_base(BLOCK_K * 128, 1024) : make_desc_base(16, 1024)) +
                          desc_addr(smem_u32(smem_b));
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = blockIdx.x; w < work_items; w += gridDim.x) {
        const int split = w / tiles;
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb0 + kb_per_split, p.num_k_blocks);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);      // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        uint32_t accum = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = a0 + (uint64_t)(stage * (C::A_BYTES >> 4));
          const uint64_t db = b0 + (uint64_t)(stage * (C::B_BYTES >> 4));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            tc_mma_bf16(d_tmem, da + k * KSTEP_A, db + k * KSTEP_B, idesc, accum);
            accum = 1;
          }
          tc_commit(&empty_bar[stage]);                   // smem slot reusable once these MMAs retire
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tmem_full[acc]);                       // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ============================ epilogue warps ============================
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    const int half = (warp - 2) >> 2;             // which half of the columns this warp drains
    const int c_begin = (BN >= 128) ? half * (BN / 2) : 0;
    const int c_end = (BN >= 128) ? c_begin + BN / 2 : (half == 0 ? BN : 0);
    uint8_t* my_store = smem_store + (warp - 2) * (2 * 4096);
    float* my_stats = s_stats + (warp - 2) * STATS_WARP_FLOATS;
    int stats_n = -1;                             // column block the shared statistics belong to
    for (int w = blockIdx.x; w < work_items; w += gridDim.x) {
      const int tile = w % tiles;
      const int m_idx = (p.n_fastest ? tile / p.num_n_blocks : tile % p.num_m_blocks) * BLOCK_M;
      const int n_idx = (p.n_fastest ? tile % p.num_n_blocks : tile / p.num_m_blocks) * BN;
      if (want_stats && n_idx != stats_n) {
        if (stats_n >= 0) stats_flush<BN>(p, s_stats, stats_n, (warp - 2) * 32 + lane);
        stats_n = n_idx;
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      epilogue_rows<BN>(p, &map_c, &map_z, tmem_base, acc, q, lane, m_idx + q * 32, n_idx, c_begin, c_end,
                        my_store, StoreAt{0, 0, 0, 0, nullptr}, want_stats ? my_stats : nullptr);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (want_stats && stats_n >= 0) stats_flush<BN>(p, s_stats, stats_n, (warp - 2) * 32 + lane);
    if (p.tma_store && lane == 0) tma_store_wait_all();   // smem must outlive the bulk reads
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}

// ====================================================================================================
// 2-CTA variant (cta_group::2): a cluster of two CTAs on one TPC computes a 256 x 256 tile.  Each CTA
// owns 128 rows of A and of the accumulator (its own TMEM) and HALF of the B tile (128 of the 256 N
// rows); the leader's single elected lane issues tcgen05.mma.cta_group::2 (UMMA 256x256x16) which
// reads A from both CTAs' smem and B halves from both — halving B traffic per CTA.  All TMA loads
// of both CTAs complete on the LEADER's full barrier; tcgen05.commit multicasts the "stage free" and
// "accumulator ready" signals to both CTAs; both CTAs' epilogue warps release the accumulator on
// the leader's tmem_empty barrier.  Arithmetic intensity per CTA: 32 KB of operands per 128x256x64
// MACs (48 KB in the 1-CTA kernel) — the L2 -> SM path is what bounds large-K GEMMs on B200.
// ====================================================================================================
constexpr int BN2 = 256;
struct Cfg2 {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;          // 16 KB : this CTA's 128 rows
  static constexpr int B_BYTES = (BN2 / 2) * BLOCK_K * 2;        // 16 KB : this CTA's half of B
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = 4;
  static constexpr int TMEM_COLS = 2 * BN2;
  static constexpr int STORE_BYTES = EPI_WARPS * 2 * 4096;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STORE_BYTES + 1024 + 256 + STATS_SMEM_BYTES;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the LEADER CTA's copy of a barrier (bit 24 selects the CTA of a pair)
__device__ __forceinline__ uint32_t leader_addr(const void* p) { return smem_u32(p) & 0xFEFFFFFFu; }

__device__ __forceinline__ void tma_load_2d_2cta(const CUtensorMap* map, uint32_t leader_bar, void* dst, int c0,
                                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], "
      "[%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2cta(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                                 uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_2cta_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_z,
                      const GemmParams p) {
  using C = Cfg2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::STAGES * C::A_BYTES;
  uint8_t* smem_store = smem + C::STAGES * C::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_store + C::STORE_BYTES);
  uint64_t* full_bar = bars;                     // [STAGES]   (used in the leader CTA only)
  uint64_t* empty_bar = bars + C::STAGES;        // [STAGES]   (per CTA; multicast-committed)
  uint64_t* tmem_full = bars + 2 * C::STAGES;    // [2]        (per CTA; multicast-committed)
  uint64_t* tmem_empty = tmem_full + 2;          // [2]        (leader CTA; 8 arrivals = 2 CTAs x 4 warps)
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_stats = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);
  const bool want_stats = p.stats != nullptr && p.out_mode == 0 && p.tma_store;
  if (want_stats) stats_zero(s_stats, NUM_THREADS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (p.tma_store) tma_prefetch_desc(&map_c);
    if (p.tma_store && p.preact != nullptr) tma_prefetch_desc(&map_z);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_base_slot)),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();                    // (CTA-scope barrier for the tcgen05.alloc result: racecheck does not model
                                      //  barrier.cluster as ordering the allocator's shared-memory write)
  cluster_sync_all();                 // both CTAs' barriers are initialised before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  const int num_m2 = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);     // 256-row blocks
  const int tiles = num_m2 * p.num_n_blocks;
  const int work_items = tiles * p.splits;
  const int kb_per_split = (p.num_k_blocks + p.splits - 1) / p.splits;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0) {
    // ============================ TMA producer (both CTAs) ============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = cluster_id; w < work_items; w += num_clusters) {
        const int tile = w % tiles, split = w / tiles;
        const int m_idx = (p.n_fastest ? tile / p.num_n_blocks : tile % num_m2) * (2 * BLOCK_M) +
                          (int)cta_rank * BLOCK_M;
        const int n_idx = (p.n_fastest ? tile % p.num_n_blocks : tile / num_m2) * BN2 +
                          (int)cta_rank * (BN2 / 2);
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb0 + kb_per_split, p.num_k_blocks);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t fb = leader_addr(&full_bar[stage]);
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);   // bytes of BOTH CTAs
          uint8_t* sa = smem_a + stage * C::A_BYTES;
          uint8_t* sb = smem_b + stage * C::B_BYTES;
          const int k_idx = kb * BLOCK_K;
          if (!A_MN) {
            tma_load_2d_2cta(&map_a, fb, sa, k_idx, m_idx);
          } else {
#pragma unroll
            for (int c = 0; c < BLOCK_M / 64; ++c)
              tma_load_2d_2cta(&map_a, fb, sa + c * (BLOCK_K * 128), m_idx + 64 * c, k_idx);
          }
          if (!B_MN) {
            tma_load_2d_2cta(&map_b, fb, sb, k_idx, n_idx);                   // box {64 k, 128 n}
          } else {
#pragma unroll
            for (int c = 0; c < (BN2 / 2) / 64; ++c)
              tma_load_2d_2cta(&map_b, fb, sb + c * (BLOCK_K * 128), n_idx + 64 * c, k_idx);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer (leader CTA only, one elected thread) ============================
    if (leader && elect_one()) {
      uint32_t idesc = 0;
      idesc |= 1u << 4;
      idesc |= 1u << 7;
      idesc |= 1u << 10;
      idesc |= (A_MN ? 1u : 0u) << 15;
      idesc |= (B_MN ? 1u : 0u) << 16;
      idesc |= (uint32_t)(BN2 >> 3) << 17;
      idesc |= (uint32_t)((2 * BLOCK_M) >> 4) << 24;       // UMMA M = 256 across the CTA pair
      constexpr uint32_t KSTEP_A = A_MN ? ((UMMA_K * 128) >> 4) : ((UMMA_K * 2) >> 4);
      constexpr uint32_t KSTEP_B = B_MN ? ((UMMA_K * 128) >> 4) : ((UMMA_K * 2) >> 4);
      const uint64_t a0 = (A_MN ? make_desc_base(BLOCK_K * 128, 1024) : make_desc_base(16, 1024)) +
                          desc_addr(smem_u32(smem_a));
      const uint64_t b0 = (B_MN ? make_desc_base(BLOCK_K * 128, 1024) : make_desc_base(16, 1024)) +
                          desc_addr(smem_u32(smem_b));
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = cluster_id; w < work_items; w += num_clusters) {
        const int split = w / tiles;
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb0 + kb_per_split, p.num_k_blocks);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN2);
        uint32_t accum = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = a0 + (uint64_t)(stage * (C::A_BYTES >> 4));
          const uint64_t db = b0 + (uint64_t)(stage * (C::B_BYTES >> 4));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            tc_mma_bf16_2cta(d_tmem, da + k * KSTEP_A, db + k * KSTEP_B, idesc, accum);
            accum = 1;
          }
          tc_commit_2cta(&empty_bar[stage]);                  // frees the stage in BOTH CTAs
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit_2cta(&tmem_full[acc]);                      // accumulators ready in BOTH CTAs
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ============================ epilogue warps (both CTAs, own 128 rows) ============================
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int half = (warp - 2) >> 2;
    const int c_begin = half * (BN2 / 2), c_end = c_begin + BN2 / 2;
    uint8_t* my_store = smem_store + (warp - 2) * (2 * 4096);
    float* my_stats = s_stats + (warp - 2) * STATS_WARP_FLOATS;
    int stats_n = -1;
    for (int w = cluster_id; w < work_items; w += num_clusters) {
      const int tile = w % tiles;
      const int m_idx = (p.n_fastest ? tile / p.num_n_blocks : tile % num_m2) * (2 * BLOCK_M) +
                        (int)cta_rank * BLOCK_M;
      const int n_idx = (p.n_fastest ? tile % p.num_n_blocks : tile / num_m2) * BN2;
      if (want_stats && n_idx != stats_n) {
        if (stats_n >= 0) stats_flush<BN2>(p, s_stats, stats_n, (warp - 2) * 32 + lane);
        stats_n = n_idx;
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      epilogue_rows<BN2>(p, &map_c, &map_z, tmem_base, acc, q, lane, m_idx + q * 32, n_idx, c_begin, c_end,
                         my_store, StoreAt{0, 0, 0, 0, nullptr}, want_stats ? my_stats : nullptr);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_addr(&tmem_empty[acc]));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (want_stats && stats_n >= 0) stats_flush<BN2>(p, s_stats, stats_n, (warp - 2) * 32 + lane);
    if (p.tma_store && lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();                 // the peer may still read my smem / signal my barriers until here
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
thread_local char g_err[512];
int g_num_sms = 0;

int fail(const char* msg, int code = 0) {
  snprintf(g_err, sizeof(g_err), "%s (%d)", msg, code);
  return -1;
}

int ensure_init() {
  bind_primary_context();
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult st;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &st);
  if (e != cudaSuccess || st != cudaDriverEntryPointSuccess || !fn)
    return fail("cuTensorMapEncodeTiled entry point unavailable", (int)e);
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  return 0;
}

// 2D bf16 tensor map: `rows` x `cols` (cols contiguous), row pitch `ld` elements, box {64, box_rows}.
// (the same encoding serves the loads of A/B and the 32-row bulk stores of C)
int make_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed", (int)r);
  return 0;
}

template <int BN, bool A_MN, bool B_MN>
int launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const CUtensorMap& mz,
           const GemmParams& p, int max_ctas, cudaStream_t st) {
  using C = Cfg<BN>;
  auto kern = gemm_bf16_kernel<BN, A_MN, B_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return fail(cudaGetErrorString(e), (int)e);
    attr_set = true;
  }
  const int work = p.num_m_blocks * p.num_n_blocks * p.splits;
  int grid = work < g_num_sms ? work : g_num_sms;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  kern<<<grid, NUM_THREADS, C::SMEM_BYTES, st>>>(ma, mb, mc, mz, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(cudaGetErrorString(e), (int)e);
  return 0;
}

template <bool A_MN, bool B_MN>
int launch2(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const CUtensorMap& mz,
            const GemmParams& p, int max_ctas, cudaStream_t st) {
  auto kern = gemm_bf16_2cta_kernel<A_MN, B_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2::SMEM_BYTES);
    if (e != cudaSuccess) return fail(cudaGetErrorString(e), (int)e);
    attr_set = true;
  }
  const int num_m2 = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int work = num_m2 * p.num_n_blocks * p.splits;
  int clusters = work < g_num_sms / 2 ? work : g_num_sms / 2;
  if (max_ctas > 1 && clusters > max_ctas / 2) clusters = max_ctas / 2;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = Cfg2::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ma, mb, mc, mz, p);
  if (e != cudaSuccess) return fail(cudaGetErrorString(e), (int)e);
  return 0;
}

}  // namespace

extern "C" {

const char* b200dp_gemm_last_error() { return g_err; }

// A: K-major -> [M][lda>=K]; MN-major -> [K][lda>=M].   B: K-major -> [N][ldb>=K]; MN-major -> [K][ldb>=N].
// C: [M][ldc>=N].  out_mode 0: bf16 = act(alpha*AB + bias) + residual; 1: fp32 atomic +=; 2: fp32 store.
// Requirements: K % 8 == 0 for K-major operands, M % 8 (A) / N % 8 (B) == 0 for MN-major, N % 8 == 0,
// 16-byte aligned base pointers and leading dimensions.
int b200dp_gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc,
                     int a_mn, int b_mn, const void* bias_bf16, const void* bias_f32, const void* residual,
                     void* preact, int act, int out_mode, float alpha, int splits, int block_n, int max_ctas,
                     int two_cta, float* stats, const void* res_mask, unsigned long long stream) {
  if (ensure_init()) return -1;
  if (M <= 0 || N <= 0 || K <= 0) return fail("bad shape");
  if ((N % 8) || (lda % 8) || (ldb % 8) || (ldc % 4) || ((out_mode == 0) && (ldc % 8)))
    return fail("alignment: N, lda, ldb, ldc must be multiples of 8");
  if (((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) & 15) return fail("pointers must be 16-byte aligned");
  int BN = block_n;
  if (BN == 0) BN = (N > 128) ? 256 : (N > 64 ? 128 : 64);
  if (BN != 64 && BN != 128 && BN != 256) return fail("block_n must be 64/128/256");
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.ldc = ldc;
  p.num_m_blocks = (M + BLOCK_M - 1) / BLOCK_M;
  p.num_n_blocks = (N + BN - 1) / BN;
  p.num_k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
  p.splits = splits < 1 ? 1 : splits;
  if (p.splits > p.num_k_blocks) p.splits = p.num_k_blocks;
  if (p.splits > 1 && out_mode != 1) return fail("split-K requires out_mode=1 (fp32 atomic add)");
  {  // no empty splits
    const int per = (p.num_k_blocks + p.splits - 1) / p.splits;
    p.splits = (p.num_k_blocks + per - 1) / per;
  }
  p.act = act; p.out_mode = out_mode; p.C = C; p.bias = bias_bf16; p.bias_f32 = bias_f32;
  p.residual = residual; p.preact = preact; p.alpha = alpha;
  p.stats = stats;
  p.res_mask = reinterpret_cast<const unsigned char*>(res_mask);
  if (res_mask != nullptr && (residual == nullptr || preact != nullptr || act > 2 || out_mode != 0 || (N % 64) ||
                              ldc != N))
    return fail("res_mask: plain bf16 residual, dense rows and N % 64 == 0 required");
  if (stats != nullptr && (N > STATS_MAX_N || out_mode != 0)) return fail("stats: N <= 2048 and bf16 output required");
  p.tma_store = (out_mode == 0) ? 1 : 0;
  // B (N x K bf16) small enough to live in L2 next to the in-flight A tiles -> walk N first
  p.n_fastest = ((size_t)N * (size_t)K * 2 <= ((size_t)48 << 20)) ? 1 : 0;
  CUtensorMap ma, mb, mc, mz;
  if (a_mn ? make_map(&ma, A, K, M, lda, BLOCK_K) : make_map(&ma, A, M, K, lda, BLOCK_M)) return -1;
  if (b_mn ? make_map(&mb, B, K, N, ldb, BLOCK_K) : make_map(&mb, B, N, K, ldb, BN)) return -1;
  if (p.tma_store) {
    if (make_map(&mc, C, M, N, ldc, 32)) return -1;
  } else {
    mc = ma;   // unused
  }
  if (p.tma_store && preact != nullptr) {
    if (make_map(&mz, preact, M, N, ldc, 32)) return -1;
  } else {
    mz = mc;   // unused
  }
  cudaStream_t st = (cudaStream_t)(uintptr_t)stream;
  if (two_cta && BN == 256) {
    // B map for the pair: each CTA loads half of the 256 N rows (K-major box {64, 128})
    if (!b_mn && make_map(&mb, B, N, K, ldb, BN2 / 2)) return -1;
    if (!a_mn && !b_mn) return launch2<false, false>(ma, mb, mc, mz, p, max_ctas, st);
    if (!a_mn && b_mn) return launch2<false, true>(ma, mb, mc, mz, p, max_ctas, st);
    if (a_mn && !b_mn) return launch2<true, false>(ma, mb, mc, mz, p, max_ctas, st);
    return launch2<true, true>(ma, mb, mc, mz, p, max_ctas, st);
  }
#define DISPATCH(BNV)                                                                         \
  if (BN == BNV) {                                                                            \
    if (!a_mn && !b_mn) return launch<BNV, false, false>(ma, mb, mc, mz, p, max_ctas, st);            \
    if (!a_mn && b_mn) return launch<BNV, false, true>(ma, mb, mc, mz, p, max_ctas, st);              \
    if (a_mn && !b_mn) return launch<BNV, true, false>(ma, mb, mc, mz, p, max_ctas, st);              \
    return launch<BNV, true, true>(ma, mb, mc, mz, p, max_ctas, st);                                  \
  }
  DISPATCH(64)
  DISPATCH(128)
  DISPATCH(256)
#undef DISPATCH
  return fail("unreachable");
}

}  // extern "C"
