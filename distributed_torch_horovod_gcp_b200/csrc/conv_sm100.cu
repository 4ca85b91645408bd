// Implicit-GEMM convolution for sm_100a (NHWC bf16): forward, data gradient and weight gradient of
// 3x3 (stride 1/2, pad 1) and 1x1 (stride 1/2) convolutions on the tcgen05 mainloop of gemm_sm100.cu.
//
// No im2col matrix is ever materialised.  An NHWC activation is a 4D tensor {C, W, H, N} for the TMA
// unit; the 128 rows of a GEMM M-tile are a {bw, bh, bn} box of output pixels, and the A operand of
// filter tap (r, s) is the SAME box shifted by (s - pad, r - pad) — one cp.async.bulk.tensor.4d per
// tap and 64-channel chunk, with the zero padding produced by the TMA's out-of-bounds fill.  The smem
// image of such a box is exactly the K-major [128 rows][64 k] 128B-swizzled tile the UMMA descriptors
// of the GEMM expect, so the MMA issuer and the TMEM epilogue are shared with the GEMM.
//
//   fprop  y[p, co]  = sum_{r,s,ci} x[p + (r,s) - pad, ci] * w[co, r, s, ci]      A = x boxes, B = w (K-major)
//   dgrad  dx[p, ci] = sum_{r,s,co} dy[p + pad - (r,s), co] * w[co, r, s, ci]     A = dy boxes, B = w (MN-major)
//   wgrad  dw[co, r, s, ci] = sum_p dy[p, co] * x[p + (r,s) - pad, ci]            A = dy boxes, B = x boxes,
//          both MN-major with K = 64-pixel boxes; split-K over pixels, fp32 RED.ADD into [Cout][R*S*Cin]
//
// Stride 2 never uses strided gathers: the input (fprop/wgrad) or the output (dgrad) is addressed
// through four parity views {C, W/2, H/2, N} (base offset (ph*W + pw)*C, doubled pitches), so every
// tap is again a dense shifted box of one view; dgrad runs the four output parities as "classes"
// with 1, 2, 2 and 4 taps.  Weights are [Cout][R][S][Cin] (PyTorch channels_last), which is the
// K-major B matrix of fprop and the MN-major B matrix of dgrad without any repacking.
//
// Replaces the cuDNN convolution calls of the ResNet zoo ([DRIVER] BASELINE.json configs 1-3;
// SURVEY.md §7.1 step 9; VERDICT r1 "next round" item 1).  The reference has no convolution at all
// (/root/reference/app/torch_train.py:107-206 is the whole model).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sm100_common.cuh"

namespace {

constexpr int MAX_TAPS = 9;
constexpr int MAX_CLASSES = 4;

struct ConvTap {
  int amap;      // which activation view (parity) the tap reads
  int dw, dh;    // box shift in that view
  int wcol;      // fprop/dgrad: column of the [Cout][R*S*Cin] weight matrix; wgrad: output column
  int row_off;   // halo kernel: first row of this tap's 8 x 16 sub-view inside the halo box
};
struct ConvClass {
  int ntaps;
  int out_map;   // which output view (dgrad stride 2: parity of dx)
  // halo kernel: ONE box {64 c, gw, gh, 1} at shift (dw0, dh0) serves every tap of the class
  int amap, dw0, dh0, gw, gh;
  ConvTap taps[MAX_TAPS];
};
struct ConvParams {
  GemmParams g;
  int num_classes;
  int kc_per_tap;                 // 64-channel chunks of the reduction dimension per tap (fprop/dgrad)
  int bw, bh, bn;                 // pixel box: 128 rows of an M tile (fprop/dgrad) / 64 rows of a K block (wgrad)
  int tiles_w, tiles_h, tiles_n;  // boxes covering the (class) output pixel space
  int num_taps_total;             // wgrad: R*S
  int out_w, out_h, out_n;        // extent of the (class) output pixel space: rows beyond it are clipped / not counted
  ConvClass cls[MAX_CLASSES];
};
struct alignas(64) ConvMaps {
  CUtensorMap a[4];     // activation views read with tap shifts (x for fprop/wgrad, dy for dgrad)
  CUtensorMap b;        // 2D weight matrix (fprop/dgrad); unused by wgrad
  CUtensorMap out[4];   // fprop/dgrad: output views (32-row slab boxes); wgrad: out[0] = dy (64-pixel boxes)
};

enum { MODE_FPROP = 0, MODE_DGRAD = 1, MODE_WGRAD = 2 };

template <int BN, int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_bf16_kernel(const __grid_constant__ ConvMaps maps, const __grid_constant__ ConvParams p) {
  using C = Cfg<BN>;
  constexpr bool A_MN = (MODE == MODE_WGRAD);
  constexpr bool B_MN = (MODE != MODE_FPROP);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::STAGES * C::A_BYTES;
  uint8_t* smem_store = smem + C::STAGES * C::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_store + C::STORE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::STAGES;
  uint64_t* tmem_full = bars + 2 * C::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_stats = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);
  const bool want_stats = MODE == MODE_FPROP && p.g.stats != nullptr;
  if (want_stats) stats_zero(s_stats, NUM_THREADS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&maps.a[i]);
    tma_prefetch_desc(&maps.b);
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&maps.out[i]);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], EPI_WARPS);
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

  const int nnb = p.g.num_n_blocks;
  const int pix_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  // fprop/dgrad: work = class x pixel tile x n block (n fastest: the A boxes of one pixel tile are
  // re-used from L2 by the CTAs working on its other n blocks).
  // wgrad: work = split x (m block x n block x tap), tap fastest: the dy / x boxes of one pixel range
  // are shared through L2 by the CTAs of the same split.
  const int tiles = (MODE == MODE_WGRAD) ? p.g.num_m_blocks * nnb * p.num_taps_total
                                         : p.num_classes * pix_tiles * nnb;
  const int work_items = tiles * p.g.splits;
  const int kb_per_split = (MODE == MODE_WGRAD) ? (pix_tiles + p.g.splits - 1) / p.g.splits : 0;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < work_items; w += gridDim.x) {
        if (MODE != MODE_WGRAD) {
          const int nt = w % nnb;
          const int rest = w / nnb;
          const int mt = rest % pix_tiles;
          const ConvClass& cl = p.cls[rest / pix_tiles];
          const int w0 = (mt % p.tiles_w) * p.bw;
          const int h0 = ((mt / p.tiles_w) % p.tiles_h) * p.bh;
          const int n0 = (mt / (p.tiles_w * p.tiles_h)) * p.bn;
          const int n_idx = nt * BN;
          for (int t = 0; t < cl.ntaps; ++t) {
            const ConvTap tap = cl.taps[t];
            for (int kc = 0; kc < p.kc_per_tap; ++kc) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
              uint8_t* sa = smem_a + stage * C::A_BYTES;
              uint8_t* sb = smem_b + stage * C::B_BYTES;
              tma_load_4d(&maps.a[tap.amap], &full_bar[stage], sa, kc * BLOCK_K, w0 + tap.dw, h0 + tap.dh, n0);
              if (!B_MN) {
                tma_load_2d(&maps.b, &full_bar[stage], sb, tap.wcol + kc * BLOCK_K, n_idx);   // box {64 k, BN co}
              } else {
#pragma unroll
                for (int c = 0; c < BN / 64; ++c)                                           // box {64 ci, 64 co}
                  tma_load_2d(&maps.b, &full_bar[stage], sb + c * (BLOCK_K * 128), tap.wcol + n_idx + 64 * c,
                              kc * BLOCK_K);
              }
              if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
          }
        } else {
          const int tile = w % tiles, split = w / tiles;
          const ConvTap tap = p.cls[0].taps[tile % p.num_taps_total];
          const int mn = tile / p.num_taps_total;
          const int m_idx = (mn / nnb) * BLOCK_M;
          const int n_idx = (mn % nnb) * BN;
          const int kb0 = split * kb_per_split;
          const int kb1 = min(kb0 + kb_per_split, pix_tiles);
          for (int kb = kb0; kb < kb1; ++kb) {
            const int w0 = (kb % p.tiles_w) * p.bw;
            const int h0 = ((kb / p.tiles_w) % p.tiles_h) * p.bh;
            const int n0 = (kb / (p.tiles_w * p.tiles_h)) * p.bn;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
            uint8_t* sa = smem_a + stage * C::A_BYTES;
            uint8_t* sb = smem_b + stage * C::B_BYTES;
#pragma unroll
            for (int c = 0; c < BLOCK_M / 64; ++c)      // dy: 64 pixels x 64 output channels per box
              tma_load_4d(&maps.out[0], &full_bar[stage], sa + c * (BLOCK_K * 128), m_idx + 64 * c, w0, h0, n0);
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)           // x : the same pixels shifted by the tap
              tma_load_4d(&maps.a[tap.amap], &full_bar[stage], sb + c * (BLOCK_K * 128), n_idx + 64 * c,
                          w0 + tap.dw, h0 + tap.dh, n0);
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
          }
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
        int nkb;
        if (MODE != MODE_WGRAD) {
          nkb = p.cls[(w / nnb) / pix_tiles].ntaps * p.kc_per_tap;
        } else {
          const int kb0 = (w / tiles) * kb_per_split;
          nkb = min(kb0 + kb_per_split, pix_tiles) - kb0;
        }
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        uint32_t accum = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = a0 + (uint64_t)(stage * (C::A_BYTES >> 4));
          const uint64_t db = b0 + (uint64_t)(stage * (C::B_BYTES >> 4));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            tc_mma_bf16(d_tmem, da + k * KSTEP_A, db + k * KSTEP_B, idesc, accum);
            accum = 1;
          }
          tc_commit(&empty_bar[stage]);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ============================ epilogue warps ============================
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int half = (warp - 2) >> 2;
    const int c_begin = (BN >= 128) ? half * (BN / 2) : 0;
    const int c_end = (BN >= 128) ? c_begin + BN / 2 : (half == 0 ? BN : 0);
    uint8_t* my_store = smem_store + (warp - 2) * (2 * 4096);
    float* my_stats = s_stats + (warp - 2) * STATS_WARP_FLOATS;
    int stats_n = -1;
    for (int w = blockIdx.x; w < work_items; w += gridDim.x) {
      if (want_stats && (w % nnb) != stats_n) {
        if (stats_n >= 0) stats_flush<BN>(p.g, s_stats, stats_n * BN, (warp - 2) * 32 + lane);
        stats_n = w % nnb;
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (MODE != MODE_WGRAD) {
        const int nt = w % nnb;
        const int rest = w / nnb;
        const int mt = rest % pix_tiles;
        const ConvClass& cl = p.cls[rest / pix_tiles];
        const int r0 = q * 32;                                   // first row of this warp's slab in the box
        StoreAt at;
        at.rank4 = 1;
        at.w = (mt % p.tiles_w) * p.bw + (r0 % p.bw);
        at.h = ((mt / p.tiles_w) % p.tiles_h) * p.bh + (r0 / p.bw) % p.bh;
        at.n = (mt / (p.tiles_w * p.tiles_h)) * p.bn + r0 / (p.bw * p.bh);
        at.c_ptr = nullptr;
        at.sw = p.bw < 32 ? p.bw : 32;
        at.sh = (32 / at.sw) < p.bh ? (32 / at.sw) : p.bh;
        at.vw = p.out_w - at.w; at.vh = p.out_h - at.h; at.vn = p.out_n - at.n;
        epilogue_rows<BN>(p.g, &maps.out[cl.out_map], nullptr, tmem_base, acc, q, lane, 0, nt * BN, c_begin,
                          c_end, my_store, at, want_stats ? my_stats : nullptr);
      } else {
        const int tile = w % tiles;
        const ConvTap tap = p.cls[0].taps[tile % p.num_taps_total];
        const int mn = tile / p.num_taps_total;
        StoreAt at;
        at.rank4 = 0; at.w = at.h = at.n = 0;
        at.sw = 32; at.sh = 1; at.vw = 32; at.vh = 1; at.vn = 1;
        at.c_ptr = reinterpret_cast<float*>(p.g.C) + tap.wcol;
        epilogue_rows<BN>(p.g, nullptr, nullptr, tmem_base, acc, q, lane, (mn / nnb) * BLOCK_M + q * 32,
                          (mn % nnb) * BN, c_begin, c_end, my_store, at);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (want_stats && stats_n >= 0) stats_flush<BN>(p.g, s_stats, stats_n * BN, (warp - 2) * 32 + lane);
    if (MODE != MODE_WGRAD && lane == 0) tma_store_wait_all();
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
// Halo variant (3x3 stride-1 fprop/dgrad, stride-2 dgrad).  The kernel above fetches the activation
// box once PER TAP, i.e. 9x from L2 — measured: it runs at the L2->SM bandwidth roofline
// (256*BN/(256+2*BN) FLOP/B: 350/650/870 TFLOP/s for BN = 64/128/256; benchmarks/conv_bench.py).
// Here an M tile is 8 (w) x 16 (h) output pixels of one image and, per 64-channel chunk, ONE TMA box
// {64 c, 8+2, 16+2, 1} brings the tile with its halo (1.4x the tile instead of 9x).  Each 8-row UMMA
// core group is one w-line of the box, so tap (r, s) is the SAME shared-memory image read through a
// descriptor whose start address is advanced by (r*gw + s) rows and whose 8-row-group stride (SBO)
// is the box line pitch gw*128 B.  The 128B swizzle is a function of the absolute smem address for
// both the TMA write and the UMMA read, which is what makes row-shifted views legal.
// Two rings: A (halo boxes, 24 KB) and B (weight tiles), because one A box feeds nine B tiles.
// ====================================================================================================
constexpr int HALO_BW = 8, HALO_BH = 16;
constexpr int HALO_A_BYTES = 24 * 1024;          // (8+2)*(16+2)*128 B = 23040, rounded to 1024
constexpr int HALO_A_STAGES = 3;
template <int BN>
struct HaloCfg {
  static constexpr int B_BYTES = BN * BLOCK_K * 2;
  static constexpr int B_STAGES = (BN == 256) ? 2 : ((BN == 128) ? 5 : 9);   // 64: all nine taps resident
  static constexpr int STORE_BYTES = EPI_WARPS * 2 * 4096;   // two staging buffers per epilogue warp
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM_BYTES = HALO_A_STAGES * HALO_A_BYTES + B_STAGES * B_BYTES + STORE_BYTES + 1024 + 256 +
                                    STATS_SMEM_BYTES;
};

template <int BN, bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_halo_kernel(const __grid_constant__ ConvMaps maps, const __grid_constant__ ConvParams p) {
  using C = HaloCfg<BN>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + HALO_A_STAGES * HALO_A_BYTES;
  uint8_t* smem_store = smem_b + C::B_STAGES * C::B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_store + C::STORE_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + HALO_A_STAGES;
  uint64_t* b_full = a_empty + HALO_A_STAGES;
  uint64_t* b_empty = b_full + C::B_STAGES;
  uint64_t* tmem_full = b_empty + C::B_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_stats = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);
  const bool want_stats = !B_MN && p.g.stats != nullptr;
  if (want_stats) stats_zero(s_stats, NUM_THREADS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&maps.a[i]);
    tma_prefetch_desc(&maps.b);
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&maps.out[i]);
    for (int i = 0; i < HALO_A_STAGES; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < C::B_STAGES; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], EPI_WARPS);
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

  const int nnb = p.g.num_n_blocks;
  const int pix_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int work_items = p.num_classes * pix_tiles * nnb;
  // Weight-stationary mode (64-channel layers): all taps of the whole filter fit in the B ring, so they are
  // loaded ONCE per CTA and the per-tap barrier wait / commit / stage bookkeeping disappears from the MMA
  // issue loop — measured, that single-thread loop (~55 instructions per tap for four N=64 MMAs), not the
  // tensor pipe (28 %) or memory (16 %), bounded the 64-channel 3x3 layers.
  const bool ws = p.kc_per_tap == 1 && nnb == 1 && p.num_classes == 1 && p.cls[0].ntaps <= C::B_STAGES;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      if (ws) {
        const ConvClass& cl = p.cls[0];
        for (int t = 0; t < cl.ntaps; ++t) {
          const int wcol = cl.taps[t].wcol;
          mbar_expect_tx(&b_full[t], C::B_BYTES);
          uint8_t* dst = smem_b + t * C::B_BYTES;
          if (!B_MN) {
            tma_load_2d(&maps.b, &b_full[t], dst, wcol, 0);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_2d(&maps.b, &b_full[t], dst + c * (BLOCK_K * 128), wcol + 64 * c, 0);
          }
        }
      }
      for (int w = blockIdx.x; w < work_items; w += gridDim.x) {
        const int nt = w % nnb;
        const int rest = w / nnb;
        const int mt = rest % pix_tiles;
        const ConvClass& cl = p.cls[rest / pix_tiles];
        const int w0 = (mt % p.tiles_w) * HALO_BW;
        const int h0 = ((mt / p.tiles_w) % p.tiles_h) * HALO_BH;
        const int n0 = mt / (p.tiles_w * p.tiles_h);
        const int n_idx = nt * BN;
        const uint32_t a_bytes = (uint32_t)cl.gw * (uint32_t)cl.gh * 128u;
        for (int kc = 0; kc < p.kc_per_tap; ++kc) {
          mbar_wait(&a_empty[sa], pa ^ 1);
          mbar_expect_tx(&a_full[sa], a_bytes);
          tma_load_4d(&maps.a[cl.amap], &a_full[sa], smem_a + sa * HALO_A_BYTES, kc * BLOCK_K, w0 + cl.dw0,
                      h0 + cl.dh0, n0);
          if (++sa == HALO_A_STAGES) { sa = 0; pa ^= 1; }
          if (ws) continue;
          for (int t = 0; t < cl.ntaps; ++t) {
            const int wcol = cl.taps[t].wcol;
            mbar_wait(&b_empty[sb], pb ^ 1);
            mbar_expect_tx(&b_full[sb], C::B_BYTES);
            uint8_t* dst = smem_b + sb * C::B_BYTES;
            if (!B_MN) {
              tma_load_2d(&maps.b, &b_full[sb], dst, wcol + kc * BLOCK_K, n_idx);
            } else {
#pragma unroll
              for (int c = 0; c < BN / 64; ++c)
                tma_load_2d(&maps.b, &b_full[sb], dst + c * (BLOCK_K * 128), wcol + n_idx + 64 * c, kc * BLOCK_K);
            }
            if (++sb == C::B_STAGES) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer (one elected thread) ============================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc<BN, false, B_MN>();
      constexpr uint32_t KSTEP_A = (UMMA_K * 2) >> 4;                            // 32 B along K (K-major)
      constexpr uint32_t KSTEP_B = B_MN ? ((UMMA_K * 128) >> 4) : ((UMMA_K * 2) >> 4);
      const uint64_t b_hi = B_MN ? make_desc_base(BLOCK_K * 128, 1024) : make_desc_base(16, 1024);
      const uint64_t b0 = b_hi + desc_addr(smem_u32(smem_b));
      const uint32_t a0 = smem_u32(smem_a);
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      if (ws) {
        const ConvClass& cl = p.cls[0];
        const int ntaps = cl.ntaps;
        const uint64_t a_hi = make_desc_base(16, (uint32_t)cl.gw * 128u);
        uint32_t row_desc[MAX_TAPS];
#pragma unroll
        for (int t = 0; t < MAX_TAPS; ++t) row_desc[t] = (uint32_t)cl.taps[t].row_off << 3;
        for (int t = 0; t < ntaps; ++t) mbar_wait(&b_full[t], 0);     // the resident filter
        tc_fence_after();
        for (int w = blockIdx.x; w < work_items; w += gridDim.x) {
          mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
          mbar_wait(&a_full[sa], pa);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
          const uint64_t a_st = a_hi + desc_addr(a0 + sa * HALO_A_BYTES);
          uint32_t accum = 0;
#pragma unroll
          for (int t = 0; t < MAX_TAPS; ++t) {
            if (t < ntaps) {
              const uint64_t da = a_st + row_desc[t];
              const uint64_t db = b0 + (uint64_t)(t * (C::B_BYTES >> 4));
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                tc_mma_bf16(d_tmem, da + k * KSTEP_A, db + k * KSTEP_B, idesc, accum);
                accum = 1;
              }
            }
          }
          tc_commit(&a_empty[sa]);
          if (++sa == HALO_A_STAGES) { sa = 0; pa ^= 1; }
          tc_commit(&tmem_full[acc]);
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      } else
      for (int w = blockIdx.x; w < work_items; w += gridDim.x) {
        const ConvClass& cl = p.cls[(w / nnb) / pix_tiles];
        const int ntaps = cl.ntaps;
        const uint64_t a_hi = make_desc_base(16, (uint32_t)cl.gw * 128u);
        uint32_t row_desc[MAX_TAPS];                                             // (row_off * 128) >> 4
#pragma unroll
        for (int t = 0; t < MAX_TAPS; ++t) row_desc[t] = (uint32_t)cl.taps[t].row_off << 3;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        uint32_t accum = 0;
        for (int kc = 0; kc < p.kc_per_tap; ++kc) {
          mbar_wait(&a_full[sa], pa);
          const uint64_t a_st = a_hi + desc_addr(a0 + sa * HALO_A_BYTES);
#pragma unroll
          for (int t = 0; t < MAX_TAPS; ++t) {
            if (t < ntaps) {
              mbar_wait(&b_full[sb], pb);
              tc_fence_after();
              const uint64_t da = a_st + row_desc[t];
              const uint64_t db = b0 + (uint64_t)(sb * (C::B_BYTES >> 4));
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                tc_mma_bf16(d_tmem, da + k * KSTEP_A, db + k * KSTEP_B, idesc, accum);
                accum = 1;
              }
              tc_commit(&b_empty[sb]);
              if (++sb == C::B_STAGES) { sb = 0; pb ^= 1; }
            }
          }
          tc_commit(&a_empty[sa]);
          if (++sa == HALO_A_STAGES) { sa = 0; pa ^= 1; }
        }
        tc_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ============================ epilogue warps ============================
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int half = (warp - 2) >> 2;
    const int c_begin = (BN >= 128) ? half * (BN / 2) : 0;
    const int c_end = (BN >= 128) ? c_begin + BN / 2 : (half == 0 ? BN : 0);
    uint8_t* my_store = smem_store + (warp - 2) * (2 * 4096);
    float* my_stats = s_stats + (warp - 2) * STATS_WARP_FLOATS;
    int stats_n = -1;
    for (int w = blockIdx.x; w < work_items; w += gridDim.x) {
      if (want_stats && (w % nnb) != stats_n) {
        if (stats_n >= 0) stats_flush<BN>(p.g, s_stats, stats_n * BN, (warp - 2) * 32 + lane);
        stats_n = w % nnb;
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int nt = w % nnb;
      const int rest = w / nnb;
      const int mt = rest % pix_tiles;
      const ConvClass& cl = p.cls[rest / pix_tiles];
      StoreAt at;
      at.rank4 = 1;                                              // slab q = lines 4q..4q+3 of the 8 x 16 tile
      at.w = (mt % p.tiles_w) * HALO_BW;
      at.h = ((mt / p.tiles_w) % p.tiles_h) * HALO_BH + 4 * q;
      at.n = mt / (p.tiles_w * p.tiles_h);
      at.c_ptr = nullptr;
      at.sw = HALO_BW; at.sh = 4;
      at.vw = p.out_w - at.w; at.vh = p.out_h - at.h; at.vn = p.out_n - at.n;
      epilogue_rows<BN>(p.g, &maps.out[cl.out_map], nullptr, tmem_base, acc, q, lane, 0, nt * BN, c_begin, c_end,
                        my_store, at, want_stats ? my_stats : nullptr);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (want_stats && stats_n >= 0) stats_flush<BN>(p.g, s_stats, stats_n * BN, (warp - 2) * 32 + lane);
    if (lane == 0) tma_store_wait_all();
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
// Weight gradient, halo variant (3x3, stride 1).  Work item = (pixel range, 64 input channels, 64 output
// channels); it owns dW[co 0..63][all 9 taps][ci 0..63] for that pixel range, all nine taps resident in
// TMEM (5 accumulators of 128 x 64: rows = {tap 2q | tap 2q+1} x 64 ci, columns = 64 co).
// Per K block of 128 pixels (8 w x 16 lines) the CTA fetches ONE x box with halo {64 ci, 10, 18} and
// ONE dy box {64 co, 8, 16}; the A operand of tap pair q is the x box seen through an MN-major
// descriptor whose start is tap 2q's row shift and whose M-chunk stride (LBO) is the row distance to
// tap 2q+1 — two taps are stacked along M without any data movement; the 8-pixel K groups are the
// box's w-lines (SBO = line pitch).  39 KB of operands feed 5 x 8 UMMAs (128x64x16), i.e. the kernel is
// tensor/smem bound instead of re-reading x nine times from L2 (plain wgrad: 24-48 KB per 128x64x64).
// Small maps: lines of consecutive images are adjacent box lines (pitch gw), and the TMA zero-fills the
// dy lines that fall outside an image, so a 7x7 map packs two images into the 16 lines of a K block.
// Results leave through coalesced fp32 RED.ADDs into the [Cout][9*Cin] workspace (rows of an
// accumulator are consecutive ci).
// ====================================================================================================
struct WgradHaloParams {
  float* ws;
  int Cin, Cout, ldc;
  int ci_chunks, co_chunks, splits;
  int tiles_w, tiles_h, tiles_n;     // K blocks: w tiles x line tiles x image groups
  int nimg;                          // images per K block (2 for maps with H + 2 <= 9)
  int line_step;                     // lines advanced per h tile (16) — one image per box when nimg == 1
  int xbytes, ybytes;                // bytes of the two boxes (expect_tx)
  int gw;                            // x box line pitch in pixels (10)
  int row_off[MAX_TAPS];             // tap row shift inside the x box
};
constexpr int WH_X_BYTES = 24 * 1024;
constexpr int WH_Y_BYTES = 18 * 1024;
constexpr int WH_STAGE_BYTES = WH_X_BYTES + WH_Y_BYTES;
constexpr int WH_STAGES = 5;
constexpr int WH_SMEM_BYTES = WH_STAGES * WH_STAGE_BYTES + 1024 + 256;
constexpr int WH_PAIRS = 5;

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_wgrad_halo_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y,
                       const __grid_constant__ WgradHaloParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WH_STAGES * WH_STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + WH_STAGES;
  uint64_t* tmem_full = bars + 2 * WH_STAGES;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_y);
    for (int i = 0; i < WH_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_base_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  const int chunk_pairs = p.ci_chunks * p.co_chunks;
  const int work_items = chunk_pairs * p.splits;
  const int kblocks = p.tiles_w * p.tiles_h * p.tiles_n;
  const int kb_per_split = (kblocks + p.splits - 1) / p.splits;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < work_items; w += gridDim.x) {
        const int cp = w % chunk_pairs, split = w / chunk_pairs;
        const int ci0 = (cp % p.ci_chunks) * 64, co0 = (cp / p.ci_chunks) * 64;
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb0 + kb_per_split, kblocks);
        for (int kb = kb0; kb < kb1; ++kb) {
          const int w0 = (kb % p.tiles_w) * 8;
          const int h0 = ((kb / p.tiles_w) % p.tiles_h) * p.line_step;
          const int n0 = (kb / (p.tiles_w * p.tiles_h)) * p.nimg;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], (uint32_t)(p.xbytes + p.ybytes));
          uint8_t* sx = smem + stage * WH_STAGE_BYTES;
          tma_load_4d(&map_x, &full_bar[stage], sx, ci0, w0 - 1, h0 - 1, n0);
          tma_load_4d(&map_y, &full_bar[stage], sx + WH_X_BYTES, co0, w0, h0, n0);
          if (++stage == WH_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // A: x box, MN-major (M = ci contiguous, K = pixels), 8-pixel K groups = box w-lines
      // B: dy box, MN-major (N = co contiguous), dense 8 x 16 pixel box
      constexpr uint32_t idesc = make_idesc<64, true, true>();
      const uint32_t line_bytes = (uint32_t)p.gw * 128u;
      uint64_t a_pair[WH_PAIRS];
#pragma unroll
      for (int q = 0; q < WH_PAIRS; ++q) {
        const int t0 = 2 * q, t1 = (2 * q + 1 < MAX_TAPS) ? 2 * q + 1 : 2 * q;
        const uint32_t lbo = (uint32_t)(p.row_off[t1] - p.row_off[t0]) * 128u;
        a_pair[q] = make_desc_base(lbo, line_bytes) + (uint64_t)((uint32_t)p.row_off[t0] << 3);
      }
      const uint64_t b_base = make_desc_base(BLOCK_K * 128, 1024);
      const uint32_t s0 = smem_u32(smem);
      const uint32_t kstep_a = (2u * line_bytes) >> 4;       // 16 pixels = two w-lines of the x box
      constexpr uint32_t kstep_b = (16 * 128) >> 4;          // 16 pixels = 16 rows of the dy box
      int stage = 0;
      uint32_t phase = 0;
      uint32_t item_phase = 0;
      for (int w = blockIdx.x; w < work_items; w += gridDim.x) {
        const int split = w / chunk_pairs;
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb0 + kb_per_split, kblocks);
        mbar_wait(tmem_empty, item_phase ^ 1);
        tc_fence_after();
        uint32_t accum = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t ax = desc_addr(s0 + stage * WH_STAGE_BYTES);
          const uint64_t by = b_base + desc_addr(s0 + stage * WH_STAGE_BYTES + WH_X_BYTES);
#pragma unroll
          for (int q = 0; q < WH_PAIRS; ++q) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              tc_mma_bf16(tmem_base + (uint32_t)(q * 64), a_pair[q] + ax + k * kstep_a, by + k * kstep_b, idesc,
                          (k > 0) ? 1u : accum);
          }
          accum = 1;
          tc_commit(&empty_bar[stage]);
          if (++stage == WH_STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(tmem_full);
        item_phase ^= 1;
      }
    }
  } else {
    // ============================ epilogue: coalesced fp32 RED.ADD ============================
    const int q4 = warp & 3;                         // TMEM lane quarter: accumulator rows 32*q4 .. +31
    const int half = (warp - 2) >> 2;                // pairs 0..2 | pairs 3..4
    uint32_t item_phase = 0;
    for (int w = blockIdx.x; w < work_items; w += gridDim.x) {
      const int cp = w % chunk_pairs;
      const int ci0 = (cp % p.ci_chunks) * 64, co0 = (cp / p.ci_chunks) * 64;
      mbar_wait(tmem_full, item_phase);
      tc_fence_after();
      const int row = q4 * 32 + lane;                // 0..127: [tap 2q | tap 2q+1] x ci
      const int ci = ci0 + (row & 63);
      for (int q = half ? 3 : 0; q < (half ? 5 : 3); ++q) {
        const int tap = 2 * q + (row >> 6);
        uint32_t r[64];
        const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(q * 64);
        tc_ld_32x32b_x32(taddr, r);
        tc_ld_32x32b_x32(taddr + 32, r + 32);
        tc_wait_ld();
        if (tap < MAX_TAPS && ci < p.Cin) {
          float* dst = p.ws + (size_t)co0 * p.ldc + tap * p.Cin + ci;
          const int nco = min(64, p.Cout - co0);
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (i < nco) atomicAdd(dst + (size_t)i * p.ldc, __uint_as_float(r[i]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);
      item_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
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

// 4D bf16 view {C, Wd, Hd, Nd} of an NHWC tensor (element pitches pw/ph/pn), box {64, bw, bh, bn}.
int make_map4(CUtensorMap* m, const void* ptr, uint64_t C, uint64_t Wd, uint64_t Hd, uint64_t Nd, uint64_t pw,
              uint64_t ph, uint64_t pn, uint32_t bw, uint32_t bh, uint32_t bn) {
  cuuint64_t dims[4] = {C, Wd, Hd, Nd};
  cuuint64_t strides[3] = {pw * 2, ph * 2, pn * 2};
  cuuint32_t box[4] = {64, bw, bh, bn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(4D) failed", (int)r);
  return 0;
}
int make_map2(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(2D) failed", (int)r);
  return 0;
}

// Pixel box {bw, bh, bn} (powers of two, bw*bh*bn == rows) that covers a W x H x N pixel space with the
// least padding; ties go to the widest box (longest contiguous runs in memory).
void choose_box(int W, int H, int N, int rows, int* bw, int* bh, int* bn) {
  double best = 1e30;
  *bw = 1; *bh = 1; *bn = rows;
  for (int w = 1; w <= rows; w <<= 1) {
    for (int h = 1; w * h <= rows; h <<= 1) {
      const int n = rows / (w * h);
      const double covered = (double)((W + w - 1) / w * w) * ((H + h - 1) / h * h) * ((N + n - 1) / n * n);
      const double score = covered - 1e-3 * w - 1e-6 * h;
      if (score < best) { best = score; *bw = w; *bh = h; *bn = n; }
    }
  }
}

int floordiv2(int e) { return (e >= 0) ? e / 2 : -((-e + 1) / 2); }

struct ConvShape {
  int N, H, W, Cin, Cout, R, S, stride, pad, OH, OW;
};

int check_shape(ConvShape& s) {
  if (s.R != s.S || (s.R != 1 && s.R != 3)) return fail("only 1x1 and 3x3 filters");
  if (s.pad != (s.R - 1) / 2) return fail("padding must be (R-1)/2");
  if (s.stride != 1 && s.stride != 2) return fail("stride must be 1 or 2");
  if (s.stride == 2 && ((s.H | s.W) & 1)) return fail("stride 2 needs even H and W");
  if ((s.Cin % 8) || (s.Cout % 8)) return fail("channels must be multiples of 8");
  s.OH = s.H / s.stride;
  s.OW = s.W / s.stride;
  return 0;
}

// parity views of an NHWC tensor with even H, W: a[ph*2+pw] = t[:, ph::2, pw::2, :]
int make_parity_maps(CUtensorMap* maps, const void* base, int C, int W, int H, int N, int bw, int bh, int bn) {
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      const __nv_bfloat16* ptr = reinterpret_cast<const __nv_bfloat16*>(base) + ((size_t)ph * W + pw) * C;
      if (make_map4(&maps[ph * 2 + pw], ptr, C, W / 2, H / 2, N, 2 * (uint64_t)C, 2 * (uint64_t)W * C,
                    (uint64_t)H * W * C, bw, bh, bn))
        return -1;
    }
  return 0;
}

void slab_box(int bw, int bh, int bn, int* sw, int* sh, int* sn) {
  *sw = bw < 32 ? bw : 32;
  *sh = (32 / *sw) < bh ? (32 / *sw) : bh;
  *sn = 32 / (*sw * *sh);
  (void)bn;
}

template <int MODE>
int launch_conv(const ConvMaps& maps, const ConvParams& p, int BN, int work, int max_ctas, cudaStream_t st) {
  int grid = work < g_num_sms ? work : g_num_sms;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
#define CONV_LAUNCH(BNV)                                                                                \
  if (BN == BNV) {                                                                                      \
    auto kern = conv_bf16_kernel<BNV, MODE>;                                                            \
    static bool attr_set = false;                                                                       \
    if (!attr_set) {                                                                                    \
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,           \
                                           Cfg<BNV>::SMEM_BYTES);                                       \
      if (e != cudaSuccess) return fail(cudaGetErrorString(e), (int)e);                                 \
      attr_set = true;                                                                                  \
    }                                                                                                   \
    kern<<<grid, NUM_THREADS, Cfg<BNV>::SMEM_BYTES, st>>>(maps, p);                                     \
    cudaError_t e = cudaGetLastError();                                                                 \
    if (e != cudaSuccess) return fail(cudaGetErrorString(e), (int)e);                                   \
    return 0;                                                                                           \
  }
  CONV_LAUNCH(64)
  CONV_LAUNCH(128)
  CONV_LAUNCH(256)
#undef CONV_LAUNCH
  return fail("block_n must be 64/128/256");
}


bool halo_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200DP_CONV_HALO");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// Re-plan a fprop/dgrad launch for the halo kernel: 8 x 16 x 1 pixel tiles, one activation box per class.
// `a_ptr` is the dense NHWC tensor {C, Wd, Hd, Nd} the taps read (x for fprop, dy for dgrad).
int setup_halo(ConvMaps& maps, ConvParams& p, const void* a_ptr, int C, int Wd, int Hd, int Nd) {
  p.bw = HALO_BW; p.bh = HALO_BH; p.bn = 1;
  p.tiles_w = (Wd + HALO_BW - 1) / HALO_BW;
  p.tiles_h = (Hd + HALO_BH - 1) / HALO_BH;
  p.tiles_n = Nd;
  for (int ci = 0; ci < p.num_classes; ++ci) {
    ConvClass& cl = p.cls[ci];
    int lo_w = 1 << 30, hi_w = -(1 << 30), lo_h = 1 << 30, hi_h = -(1 << 30);
    for (int t = 0; t < cl.ntaps; ++t) {
      lo_w = cl.taps[t].dw < lo_w ? cl.taps[t].dw : lo_w; hi_w = cl.taps[t].dw > hi_w ? cl.taps[t].dw : hi_w;
      lo_h = cl.taps[t].dh < lo_h ? cl.taps[t].dh : lo_h; hi_h = cl.taps[t].dh > hi_h ? cl.taps[t].dh : hi_h;
    }
    cl.amap = ci; cl.dw0 = lo_w; cl.dh0 = lo_h;
    cl.gw = HALO_BW + (hi_w - lo_w); cl.gh = HALO_BH + (hi_h - lo_h);
    if (cl.gw * cl.gh * 128 > HALO_A_BYTES) return fail("halo box too large");
    for (int t = 0; t < cl.ntaps; ++t)
      cl.taps[t].row_off = (cl.taps[t].dh - lo_h) * cl.gw + (cl.taps[t].dw - lo_w);
    if (make_map4(&maps.a[ci], a_ptr, C, Wd, Hd, Nd, C, (uint64_t)Wd * C, (uint64_t)Hd * Wd * C, cl.gw, cl.gh, 1))
      return -1;
  }
  for (int ci = p.num_classes; ci < 4; ++ci) maps.a[ci] = maps.a[0];
  return 0;
}

template <bool B_MN>
int launch_halo(const ConvMaps& maps, const ConvParams& p, int BN, int work, int max_ctas, cudaStream_t st) {
  int grid = work < g_num_sms ? work : g_num_sms;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
#define HALO_LAUNCH(BNV)                                                                                \
  if (BN == BNV) {                                                                                      \
    auto kern = conv_halo_kernel<BNV, B_MN>;                                                            \
    static bool attr_set = false;                                                                       \
    if (!attr_set) {                                                                                    \
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,           \
                                           HaloCfg<BNV>::SMEM_BYTES);                                   \
      if (e != cudaSuccess) return fail(cudaGetErrorString(e), (int)e);                                 \
      attr_set = true;                                                                                  \
    }                                                                                                   \
    kern<<<grid, NUM_THREADS, HaloCfg<BNV>::SMEM_BYTES, st>>>(maps, p);                                 \
    cudaError_t e = cudaGetLastError();                                                                 \
    if (e != cudaSuccess) return fail(cudaGetErrorString(e), (int)e);                                   \
    return 0;                                                                                           \
  }
  HALO_LAUNCH(64)
  HALO_LAUNCH(128)
  HALO_LAUNCH(256)
#undef HALO_LAUNCH
  return fail("block_n must be 64/128/256");
}


int launch_wgrad_halo(const void* dy, const void* x, void* dw_acc, const ConvShape& s, int splits, int max_ctas,
                      cudaStream_t st) {
  WgradHaloParams p;
  memset(&p, 0, sizeof(p));
  p.ws = reinterpret_cast<float*>(dw_acc);
  p.Cin = s.Cin; p.Cout = s.Cout; p.ldc = 9 * s.Cin;
  p.ci_chunks = (s.Cin + 63) / 64; p.co_chunks = (s.Cout + 63) / 64;
  p.gw = 10;
  p.nimg = (s.H + 2 <= 9) ? 2 : 1;
  const int xlines = (p.nimg == 2) ? (s.H + 2) : 18;      // lines of the x box per image
  const int ylines = (p.nimg == 2) ? (s.H + 2) : 16;      // dy lines per image (rows past H are zero-filled)
  p.line_step = 16;
  p.tiles_w = (s.W + 7) / 8;
  p.tiles_h = (p.nimg == 2) ? 1 : (s.H + 15) / 16;
  p.tiles_n = (s.N + p.nimg - 1) / p.nimg;
  p.xbytes = 10 * xlines * p.nimg * 128;
  p.ybytes = 8 * ylines * p.nimg * 128;
  if (p.xbytes > WH_X_BYTES || p.ybytes > WH_Y_BYTES) return fail("wgrad halo boxes too large");
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) p.row_off[r * 3 + c] = r * p.gw + c;
  CUtensorMap mx, my;
  if (make_map4(&mx, x, s.Cin, s.W, s.H, s.N, s.Cin, (uint64_t)s.W * s.Cin, (uint64_t)s.H * s.W * s.Cin, 10, xlines,
                p.nimg))
    return -1;
  if (make_map4(&my, dy, s.Cout, s.W, s.H, s.N, s.Cout, (uint64_t)s.W * s.Cout, (uint64_t)s.H * s.W * s.Cout, 8,
                ylines, p.nimg))
    return -1;
  const int kblocks = p.tiles_w * p.tiles_h * p.tiles_n;
  const int cps = p.ci_chunks * p.co_chunks;
  if (splits <= 0) {
    splits = g_num_sms / cps;                 // one wave
    if (splits < 1) splits = 1;
  }
  if (splits > kblocks) splits = kblocks;
  {
    const int per = (kblocks + splits - 1) / splits;
    splits = (kblocks + per - 1) / per;
  }
  p.splits = splits;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         WH_SMEM_BYTES);
    if (e != cudaSuccess) return fail(cudaGetErrorString(e), (int)e);
    attr_set = true;
  }
  int grid = cps * splits;
  if (grid > g_num_sms) grid = g_num_sms;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  conv_wgrad_halo_kernel<<<grid, NUM_THREADS, WH_SMEM_BYTES, st>>>(mx, my, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(cudaGetErrorString(e), (int)e);
  return 0;
}

void init_gemm_params(GemmParams& g) {
  memset(&g, 0, sizeof(g));
  g.M = INT_MAX;
  g.splits = 1;
  g.alpha = 1.0f;
  g.n_fastest = 1;
}

int pick_bn(int n, int block_n) {
  if (block_n) return block_n;
  return (n > 128) ? 256 : (n > 64 ? 128 : 64);
}

}  // namespace

extern "C" {

const char* b200dp_conv_last_error() { return g_err; }

// y[N, OH, OW, Cout] = conv(x[N, H, W, Cin], w[Cout, R, S, Cin])          (all bf16, NHWC / KRSC)
int b200dp_conv_fprop(const void* x, const void* w, void* y, int N, int H, int W, int Cin, int Cout, int R, int S,
                      int stride, int pad, int block_n, int max_ctas, float* stats, unsigned long long stream) {
  if (ensure_init()) return -1;
  ConvShape s{N, H, W, Cin, Cout, R, S, stride, pad, 0, 0};
  if (check_shape(s)) return -1;
  if (((uintptr_t)x | (uintptr_t)w | (uintptr_t)y) & 15) return fail("pointers must be 16-byte aligned");
  const int BN = pick_bn(Cout, block_n);
  ConvMaps maps;
  ConvParams p;
  memset(&p, 0, sizeof(p));
  init_gemm_params(p.g);
  p.g.N = Cout; p.g.K = Cin; p.g.ldc = Cout; p.g.C = y; p.g.out_mode = 0; p.g.tma_store = 1;
  p.g.stats = stats;
  if (stats != nullptr && Cout > STATS_MAX_N) return fail("stats: Cout <= 2048 required");
  p.g.num_n_blocks = (Cout + BN - 1) / BN;
  p.g.num_k_blocks = (Cin + BLOCK_K - 1) / BLOCK_K;
  p.kc_per_tap = p.g.num_k_blocks;
  choose_box(s.OW, s.OH, N, BLOCK_M, &p.bw, &p.bh, &p.bn);
  p.tiles_w = (s.OW + p.bw - 1) / p.bw; p.tiles_h = (s.OH + p.bh - 1) / p.bh; p.tiles_n = (N + p.bn - 1) / p.bn;
  p.num_classes = 1;
  p.num_taps_total = R * S;
  p.out_w = s.OW; p.out_h = s.OH; p.out_n = N;
  ConvClass& cl = p.cls[0];
  cl.out_map = 0;
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < S; ++c) {
      ConvTap& t = cl.taps[cl.ntaps++];
      const int eh = r - pad, ew = c - pad;
      if (stride == 1) {
        t.amap = 0; t.dh = eh; t.dw = ew;
      } else {
        const int ph = eh & 1, pw = ew & 1;
        t.amap = ph * 2 + pw; t.dh = floordiv2(eh - ph); t.dw = floordiv2(ew - pw);
      }
      t.wcol = (r * S + c) * Cin;
    }
  // measured (benchmarks/conv_bench.py): the halo tile's 8 x 16 shape wastes up to 30 % of the MMAs on 28/14-pixel
  // maps, which only pays while the kernel is operand-fetch bound, i.e. for narrow N tiles
  const bool halo = halo_enabled() && R == 3 && stride == 1 && s.OH >= 12 && s.OW >= 8 && Cout <= 128;
  if (halo) {
    if (setup_halo(maps, p, x, Cin, W, H, N)) return -1;
  } else if (stride == 1) {
    if (make_map4(&maps.a[0], x, Cin, W, H, N, Cin, (uint64_t)W * Cin, (uint64_t)H * W * Cin, p.bw, p.bh, p.bn))
      return -1;
    maps.a[1] = maps.a[2] = maps.a[3] = maps.a[0];
  } else if (make_parity_maps(maps.a, x, Cin, W, H, N, p.bw, p.bh, p.bn)) {
    return -1;
  }
  if (make_map2(&maps.b, w, Cout, (uint64_t)R * S * Cin, (uint64_t)R * S * Cin, BN)) return -1;
  int sw, sh, sn;
  slab_box(p.bw, p.bh, p.bn, &sw, &sh, &sn);
  if (make_map4(&maps.out[0], y, Cout, s.OW, s.OH, N, Cout, (uint64_t)s.OW * Cout, (uint64_t)s.OH * s.OW * Cout, sw,
                sh, sn))
    return -1;
  maps.out[1] = maps.out[2] = maps.out[3] = maps.out[0];
  const int work = p.tiles_w * p.tiles_h * p.tiles_n * p.g.num_n_blocks;
  if (halo) return launch_halo<false>(maps, p, BN, work, max_ctas, (cudaStream_t)(uintptr_t)stream);
  return launch_conv<MODE_FPROP>(maps, p, BN, work, max_ctas, (cudaStream_t)(uintptr_t)stream);
}

// dx[N, H, W, Cin] = conv_transpose(dy[N, OH, OW, Cout], w[Cout, R, S, Cin])
int b200dp_conv_dgrad(const void* dy, const void* w, void* dx, int N, int H, int W, int Cin, int Cout, int R, int S,
                      int stride, int pad, int block_n, int max_ctas, unsigned long long stream) {
  if (ensure_init()) return -1;
  ConvShape s{N, H, W, Cin, Cout, R, S, stride, pad, 0, 0};
  if (check_shape(s)) return -1;
  if (((uintptr_t)dy | (uintptr_t)w | (uintptr_t)dx) & 15) return fail("pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)(uintptr_t)stream;
  const int BN = pick_bn(Cin, block_n);
  ConvMaps maps;
  ConvParams p;
  memset(&p, 0, sizeof(p));
  init_gemm_params(p.g);
  p.g.N = Cin; p.g.K = Cout; p.g.ldc = Cin; p.g.C = dx; p.g.out_mode = 0; p.g.tma_store = 1;
  p.g.num_n_blocks = (Cin + BN - 1) / BN;
  p.g.num_k_blocks = (Cout + BLOCK_K - 1) / BLOCK_K;
  p.kc_per_tap = p.g.num_k_blocks;
  choose_box(s.OW, s.OH, N, BLOCK_M, &p.bw, &p.bh, &p.bn);   // stride 2: each dx parity view is OW x OH
  p.tiles_w = (s.OW + p.bw - 1) / p.bw; p.tiles_h = (s.OH + p.bh - 1) / p.bh; p.tiles_n = (N + p.bn - 1) / p.bn;
  p.num_taps_total = R * S;
  p.out_w = s.OW; p.out_h = s.OH; p.out_n = N;
  const bool halo = halo_enabled() && R == 3 && stride == 1 && s.OH >= 12 && s.OW >= 8 && Cin <= 128;
  if (halo) { p.bw = HALO_BW; p.bh = HALO_BH; p.bn = 1; }
  int sw, sh, sn;
  slab_box(p.bw, p.bh, p.bn, &sw, &sh, &sn);
  if (!halo) {
    if (make_map4(&maps.a[0], dy, Cout, s.OW, s.OH, N, Cout, (uint64_t)s.OW * Cout, (uint64_t)s.OH * s.OW * Cout,
                  p.bw, p.bh, p.bn))
      return -1;
    maps.a[1] = maps.a[2] = maps.a[3] = maps.a[0];
  }
  if (make_map2(&maps.b, w, Cout, (uint64_t)R * S * Cin, (uint64_t)R * S * Cin, 64)) return -1;
  if (stride == 1) {
    p.num_classes = 1;
    ConvClass& cl = p.cls[0];
    cl.out_map = 0;
    for (int r = 0; r < R; ++r)
      for (int c = 0; c < S; ++c) {
        ConvTap& t = cl.taps[cl.ntaps++];
        t.amap = 0; t.dh = pad - r; t.dw = pad - c; t.wcol = (r * S + c) * Cin;
      }
    if (make_map4(&maps.out[0], dx, Cin, W, H, N, Cin, (uint64_t)W * Cin, (uint64_t)H * W * Cin, sw, sh, sn))
      return -1;
    maps.out[1] = maps.out[2] = maps.out[3] = maps.out[0];
  } else {
    if (make_parity_maps(maps.out, dx, Cin, W, H, N, sw, sh, sn)) return -1;
    bool any_empty = false;
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        ConvClass cl;
        memset(&cl, 0, sizeof(cl));
        cl.out_map = ph * 2 + pw;
        for (int r = 0; r < R; ++r) {
          if ((ph + pad - r) & 1) continue;
          for (int c = 0; c < S; ++c) {
            if ((pw + pad - c) & 1) continue;
            ConvTap& t = cl.taps[cl.ntaps++];
            t.amap = 0; t.dh = floordiv2(ph + pad - r); t.dw = floordiv2(pw + pad - c); t.wcol = (r * S + c) * Cin;
          }
        }
        if (cl.ntaps == 0) { any_empty = true; continue; }
        p.cls[p.num_classes++] = cl;
      }
    if (any_empty) {    // 1x1 stride 2: the odd rows / columns of dx receive no gradient
      cudaError_t e = cudaMemsetAsync(dx, 0, (size_t)N * H * W * Cin * 2, st);
      if (e != cudaSuccess) return fail(cudaGetErrorString(e), (int)e);
    }
  }
  if (halo && setup_halo(maps, p, dy, Cout, s.OW, s.OH, N)) return -1;
  const int work = p.num_classes * p.tiles_w * p.tiles_h * p.tiles_n * p.g.num_n_blocks;
  if (halo) return launch_halo<true>(maps, p, BN, work, max_ctas, st);
  return launch_conv<MODE_DGRAD>(maps, p, BN, work, max_ctas, st);
}

// dw_acc[Cout][R*S*Cin] (fp32) += dy^T (*) x      — split-K over pixels with RED.ADD; the caller zeroes
// dw_acc before the first call and converts it to the weight dtype afterwards.
int b200dp_conv_wgrad(const void* dy, const void* x, void* dw_acc, int N, int H, int W, int Cin, int Cout, int R,
                      int S, int stride, int pad, int splits, int block_n, int max_ctas, unsigned long long stream) {
  if (ensure_init()) return -1;
  ConvShape s{N, H, W, Cin, Cout, R, S, stride, pad, 0, 0};
  if (check_shape(s)) return -1;
  if (((uintptr_t)dy | (uintptr_t)x | (uintptr_t)dw_acc) & 15) return fail("pointers must be 16-byte aligned");
  if (halo_enabled() && R == 3 && stride == 1 && W >= 7 && H >= 7 && block_n == 0 && (long)Cin * Cout <= 65536)
    return launch_wgrad_halo(dy, x, dw_acc, s, splits, max_ctas, (cudaStream_t)(uintptr_t)stream);
  const int BN = pick_bn(Cin, block_n);
  ConvMaps maps;
  ConvParams p;
  memset(&p, 0, sizeof(p));
  init_gemm_params(p.g);
  p.g.M = Cout; p.g.N = Cin; p.g.ldc = R * S * Cin; p.g.C = dw_acc; p.g.out_mode = 1; p.g.tma_store = 0;
  p.g.num_m_blocks = (Cout + BLOCK_M - 1) / BLOCK_M;
  p.g.num_n_blocks = (Cin + BN - 1) / BN;
  choose_box(s.OW, s.OH, N, BLOCK_K, &p.bw, &p.bh, &p.bn);     // 64-pixel K blocks
  p.tiles_w = (s.OW + p.bw - 1) / p.bw; p.tiles_h = (s.OH + p.bh - 1) / p.bh; p.tiles_n = (N + p.bn - 1) / p.bn;
  const int kblocks = p.tiles_w * p.tiles_h * p.tiles_n;
  p.num_classes = 1;
  p.num_taps_total = R * S;
  ConvClass& cl = p.cls[0];
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < S; ++c) {
      ConvTap& t = cl.taps[cl.ntaps++];
      const int eh = r - pad, ew = c - pad;
      if (stride == 1) {
        t.amap = 0; t.dh = eh; t.dw = ew;
      } else {
        const int ph = eh & 1, pw = ew & 1;
        t.amap = ph * 2 + pw; t.dh = floordiv2(eh - ph); t.dw = floordiv2(ew - pw);
      }
      t.wcol = (r * S + c) * Cin;
    }
  const int tiles = p.g.num_m_blocks * p.g.num_n_blocks * R * S;
  if (splits <= 0) {
    splits = g_num_sms / tiles;                          // one wave (148 SMs): never a second, short one
    if (splits < 1) splits = 1;
    const int max_splits = kblocks / 8 > 1 ? kblocks / 8 : 1;
    if (splits > max_splits) splits = max_splits;
  }
  if (splits > kblocks) splits = kblocks;
  if (splits < 1) splits = 1;
  {  // no empty splits
    const int per = (kblocks + splits - 1) / splits;
    splits = (kblocks + per - 1) / per;
  }
  p.g.splits = splits;
  if (stride == 1) {
    if (make_map4(&maps.a[0], x, Cin, W, H, N, Cin, (uint64_t)W * Cin, (uint64_t)H * W * Cin, p.bw, p.bh, p.bn))
      return -1;
    maps.a[1] = maps.a[2] = maps.a[3] = maps.a[0];
  } else if (make_parity_maps(maps.a, x, Cin, W, H, N, p.bw, p.bh, p.bn)) {
    return -1;
  }
  if (make_map4(&maps.out[0], dy, Cout, s.OW, s.OH, N, Cout, (uint64_t)s.OW * Cout, (uint64_t)s.OH * s.OW * Cout,
                p.bw, p.bh, p.bn))
    return -1;
  maps.out[1] = maps.out[2] = maps.out[3] = maps.out[0];
  maps.b = maps.out[0];   // unused
  return launch_conv<MODE_WGRAD>(maps, p, BN, tiles * splits, max_ctas, (cudaStream_t)(uintptr_t)stream);
}

}  // extern "C"
