// Flash attention for sm_100a (bf16, head dim 64, non-causal): forward and backward on tcgen05.
//
//   O = softmax(Q K^T / sqrt(d)) V           per (batch, head); S x S scores never leave the SM.
//
// Forward, one CTA per (batch, head, 128-query tile), two CTAs resident per SM:
//   warp 4     TMA producer: Q tile once, then K_0, V_0, K_1, V_1, ... through a 3-tile ring
//              (4D tensor maps {64 d, S, H, B}: any [B,H,S,d] / [B,S,H,d] strided layout, rows >= S zero-filled)
//   warp 5     one elected thread issues  S_j = Q K_j^T  (UMMA 128 x N_j x 16, fp32 in TMEM) and
//              O_j = P_j V_j (UMMA 128 x 64 x 16, V as the MN-major operand straight from its [key][d] tile)
//   warps 0-3  softmax: thread == query row == TMEM lane.  Two passes over the S_j columns with tcgen05.ld
//              (row max, then exp2 / row sum), P_j written as bf16 into the K-major 128B-swizzled layout the
//              PV UMMA reads; the running output lives in registers (O += P_j V_j read back from TMEM,
//              rescaled by 2^(m_old - m_new)), so TMEM needs no read-modify-write.
// The softmax is exp-bound (MUFU), the MMAs and TMA of one CTA hide behind the softmax of the other.
//
// Backward, one CTA per (batch, head, 128-key block) looping over the query tiles (FlashAttention-2
// order): S = Q_i K_j^T and dP = dO_i V_j^T on the tensor core; P = exp2(S c - LSE), dS = P (dP - D) / sqrt(d)
// in registers -> bf16 smem; dV_j += P^T dO_i, dK_j += dS^T Q_i (A operands = the SAME smem tiles read
// MN-major, no transposes), dQ_i = dS K_j -> fp32 RED.ADD into a workspace (the only cross-CTA reduction).
//
// Replaces F.scaled_dot_product_attention (cuDNN / flash library kernels) on the ViT-B/16 path
// ([DRIVER] BASELINE.json config 4; VERDICT r1 missing item 3).  The reference has
// no attention (/root/reference/app/torch_train.py:1-312).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "sm100_common.cuh"

namespace {

constexpr int AT = 192;              // threads: warps 0-3 softmax, 4 TMA, 5 MMA
constexpr int HD = 64;               // head dim
constexpr int TILE = 128;            // query rows / keys per block
constexpr int TILE_BYTES = TILE * HD * 2;   // 16 KB

struct AttnFwdParams {
  int B, H, S, q_tiles, kv_blocks;
  float scale_log2;                  // log2(e) / sqrt(d)
  __nv_bfloat16* o;
  long long o_sb, o_sh, o_ss;        // element strides of O (d contiguous)
  float* lse;                        // [B][H][S] natural-log sum-exp of the scaled scores (nullptr: not saved)
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cnt(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_cta() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ uint32_t idesc_bf16(int n, bool a_mn, bool b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 1u << 7;
  d |= 1u << 10;
  d |= (a_mn ? 1u : 0u) << 15;
  d |= (b_mn ? 1u : 0u) << 16;
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(128 >> 4) << 24;
  return d;
}

// ---------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------
constexpr int FW_RING = 3;
constexpr int FW_SMEM = TILE_BYTES * (1 + FW_RING) + 2 * TILE_BYTES /*P*/ + 1024 + 256;

__global__ void __launch_bounds__(AT, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                const __grid_constant__ CUtensorMap map_v, const AttnFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sT = smem + TILE_BYTES;                       // ring of K / V tiles
  uint8_t* sP = sT + FW_RING * TILE_BYTES;               // 2 chunks (64 keys each) x [128 rows x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                          // [3]
  uint64_t* kv_empty = bars + 4;                         // [3]
  uint64_t* s_full = bars + 7;
  uint64_t* p_full = bars + 8;
  uint64_t* o_full = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x % p.q_tiles;
  const int bh = blockIdx.x / p.q_tiles;
  const int h = bh % p.H, b = bh / p.H;
  const int nb = p.kv_blocks;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < FW_RING; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_s = tmem, tmem_o = tmem + 128;

  if (warp == 4) {
    if (elect_one()) {
      mbar_expect_tx(q_full, TILE_BYTES);
      tma_load_4d(&map_q, q_full, sQ, 0, qt * TILE, h, b);
      for (int t = 0; t < 2 * nb; ++t) {
        const int st = t % FW_RING;
        mbar_wait(&kv_empty[st], ((t / FW_RING) & 1) ^ 1);
        mbar_expect_tx(&kv_full[st], TILE_BYTES);
        tma_load_4d((t & 1) ? &map_v : &map_k, &kv_full[st], sT + st * TILE_BYTES, 0, (t >> 1) * TILE, h, b);
      }
    }
  } else if (warp == 5) {
    if (elect_one()) {
      const uint64_t kmaj = make_desc_base(16, 1024);
      const uint64_t mnmaj = make_desc_base(BLOCK_K * 128, 1024);
      const uint64_t dq = kmaj + desc_addr(smem_u32(sQ));
      const uint32_t sT_u = smem_u32(sT), sP_u = smem_u32(sP);
      mbar_wait(q_full, 0);
      int t = 0;
      auto issue_s = [&](int j) {
        const int st = t % FW_RING;
        const int valid = min(TILE, p.S - j * TILE);
        const int npad = (valid + 15) & ~15;
        mbar_wait(&kv_full[st], (t / FW_RING) & 1);
        tc_fence_after();
        const uint64_t dk = kmaj + desc_addr(sT_u + st * TILE_BYTES);
        const uint32_t id = idesc_bf16(npad, false, false);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) tc_mma_bf16(tmem_s, dq + 2 * k, dk + 2 * k, id, k ? 1u : 0u);
        tc_commit(&kv_empty[st]);
        tc_commit(s_full);
        ++t;
      };
      issue_s(0);
      for (int j = 0; j < nb; ++j) {
        const int valid = min(TILE, p.S - j * TILE);
        const int npad = (valid + 15) & ~15;
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const int st = t % FW_RING;
        mbar_wait(&kv_full[st], (t / FW_RING) & 1);
        tc_fence_after();
        const uint64_t dv = mnmaj + desc_addr(sT_u + st * TILE_BYTES);
        const uint32_t id = idesc_bf16(HD, false, true);
        for (int k = 0; k < npad / 16; ++k) {
          const uint64_t dp = kmaj + desc_addr(sP_u + (k >> 2) * TILE_BYTES + (k & 3) * 32);
          tc_mma_bf16(tmem_o, dp, dv + (uint64_t)(k * 128), id, k ? 1u : 0u);
        }
        tc_commit(&kv_empty[st]);
        tc_commit(o_full);
        ++t;
        if (j + 1 < nb) issue_s(j + 1);
      }
    }
  } else {
    // ============================ softmax warps ============================
    const int row = warp * 32 + lane;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t sP_u = smem_u32(sP);
    float m = -INFINITY, l = 0.f;
    float acc[HD];
#pragma unroll
    for (int i = 0; i < HD; ++i) acc[i] = 0.f;
    for (int j = 0; j < nb; ++j) {
      const int valid = min(TILE, p.S - j * TILE);
      const int npad = (valid + 15) & ~15;
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // ---- pass 1: row max over the valid columns
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c * 32 < npad) {
          uint32_t r[32];
          tc_ld_32x32b_x32(tmem_s + lane_base + c * 32, r);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
      }
      const float m_new = fmaxf(m, mx * p.scale_log2);
      const float alpha = (m == -INFINITY) ? 0.f : ex2(m - m_new);
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1);        // O_{j-1} = P_{j-1} V_{j-1} has landed (and sP is free again)
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tc_ld_32x32b_x32(tmem_o + lane_base + c * 32, r);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[c * 32 + i] += __uint_as_float(r[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < HD; ++i) acc[i] *= alpha;
      l *= alpha;
      m = m_new;
      // ---- pass 2: P = 2^(s c - m), row sum, bf16 -> swizzled K-major smem
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c * 32 < npad) {
          uint32_t r[32];
          tc_ld_32x32b_x32(tmem_s + lane_base + c * 32, r);
          tc_wait_ld();
          float pv[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float e = ex2(fmaf(__uint_as_float(r[i]), p.scale_log2, -m_new));
            pv[i] = (c * 32 + i < valid) ? e : 0.f;
            l += pv[i];
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int u = (c & 1) * 4 + q;                     // 16-byte unit inside the 128-byte row
            sts128(sP_u + (c >> 1) * TILE_BYTES + row * 128 + ((u ^ (row & 7)) << 4), pack8(pv + q * 8));
          }
        }
      }
      tc_fence_before();
      fence_async_cta();
      mbar_arrive_cnt(p_full);
    }
    mbar_wait(o_full, (nb - 1) & 1);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tc_ld_32x32b_x32(tmem_o + lane_base + c * 32, r);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[c * 32 + i] += __uint_as_float(r[i]);
    }
    tc_fence_before();
    const int qrow = qt * TILE + row;
    if (qrow < p.S) {
      const float inv = __fdividef(1.0f, l);
#pragma unroll
      for (int i = 0; i < HD; ++i) acc[i] *= inv;
      uint4* dst = reinterpret_cast<uint4*>(p.o + (size_t)b * p.o_sb + (size_t)h * p.o_sh + (size_t)qrow * p.o_ss);
#pragma unroll
      for (int q = 0; q < 8; ++q) dst[q] = pack8(acc + q * 8);
      if (p.lse != nullptr) p.lse[((size_t)b * p.H + h) * p.S + qrow] = (m + log2f(l)) * 0.6931471805599453f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------
struct AttnBwdParams {
  int B, H, S, q_tiles, kv_blocks;
  float scale_log2, scale;
  const float* lse;                  // [B][H][S]
  const float* delta;                // [B][H][S]  rowsum(dO * O)
  float* dq_acc;                     // fp32, zero on entry; element strides dq_sb / dq_sh / dq_ss (64 contiguous)
  long long dq_sb, dq_sh, dq_ss;
  __nv_bfloat16* dk; __nv_bfloat16* dv;
  long long dk_sb, dk_sh, dk_ss, dv_sb, dv_sh, dv_ss;
};

// delta[b][h][s] = sum_d dO * O   (8 lanes per row: one 16-byte load of each tensor per lane)
__global__ void __launch_bounds__(256) attn_delta_kernel(const __nv_bfloat16* __restrict__ o,
                                                         const __nv_bfloat16* __restrict__ dout,
                                                         float* __restrict__ delta, int B, int H, int S, long long o_sb,
                                                         long long o_sh, long long o_ss, long long d_sb, long long d_sh,
                                                         long long d_ss) {
  const unsigned total = (unsigned)B * (unsigned)H * (unsigned)S;
  const unsigned row = blockIdx.x * 32u + (threadIdx.x >> 3);
  const int sub = threadIdx.x & 7;
  float v = 0.f;
  if (row < total) {
    const unsigned s = row % (unsigned)S;
    const unsigned bh = row / (unsigned)S;
    const unsigned h = bh % (unsigned)H, b = bh / (unsigned)H;
    const uint4 ov = *reinterpret_cast<const uint4*>(o + b * o_sb + h * o_sh + s * o_ss + sub * 8);
    const uint4 dv = *reinterpret_cast<const uint4*>(dout + b * d_sb + h * d_sh + s * d_ss + sub * 8);
    float a[8], c[8];
    unpack8(ov, a);
    unpack8(dv, c);
#pragma unroll
    for (int i = 0; i < 8; ++i) v = fmaf(a[i], c[i], v);
  }
#pragma unroll
  for (int off = 4; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  if (sub == 0 && row < total) delta[row] = v;
}

// smem: K_j, V_j (resident), ring of 2 x {Q_i, dO_i}, P (32 KB), dS (32 KB)
constexpr int BW_RING = 2;
constexpr int BW_SMEM = TILE_BYTES * (2 + 2 * BW_RING) + 4 * TILE_BYTES + 1024 + 256;
// TMEM columns: S 0..127 | dP 128..255 | dV 256..319 | dK 320..383 | dQ 384..447
constexpr uint32_t TM_S = 0, TM_DP = 128, TM_DV = 256, TM_DK = 320, TM_DQ = 384;

__global__ void __launch_bounds__(AT, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_do,
                const AttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sK = smem;
  uint8_t* sV = smem + TILE_BYTES;
  uint8_t* sR = sV + TILE_BYTES;                         // ring: [stage][Q | dO]
  uint8_t* sP = sR + 2 * BW_RING * TILE_BYTES;           // P  [128 q rows][128 keys] bf16, 2 chunks
  uint8_t* sS = sP + 2 * TILE_BYTES;                     // dS same layout
  uint64_t* bars = reinterpret_cast<uint64_t*>(sS + 2 * TILE_BYTES);
  uint64_t* kv_full = bars;
  uint64_t* r_full = bars + 1;                           // [2]
  uint64_t* r_empty = bars + 3;                          // [2]
  uint64_t* sdp_full = bars + 5;                         // S and dP accumulators ready
  uint64_t* pds_full = bars + 6;                         // P and dS written to smem (128 arrivals)
  uint64_t* dq_full = bars + 7;                          // dQ_i accumulator ready (also: P / dS smem free again)
  uint64_t* dq_empty = bars + 8;                         // dQ_i drained from TMEM (128 arrivals)
  uint64_t* fin_full = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb = blockIdx.x % p.kv_blocks;
  const int bh = blockIdx.x / p.kv_blocks;
  const int h = bh % p.H, b = bh / p.H;
  const int nq = p.q_tiles;
  const int kvalid = min(TILE, p.S - kb * TILE);
  const int kpad = (kvalid + 15) & ~15;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
    tma_prefetch_desc(&map_do);
    mbar_init(kv_full, 1);
    for (int i = 0; i < BW_RING; ++i) {
      mbar_init(&r_full[i], 1);
      mbar_init(&r_empty[i], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(pds_full, 128);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 128);
    mbar_init(fin_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    if (elect_one()) {
      mbar_expect_tx(kv_full, 2 * TILE_BYTES);
      tma_load_4d(&map_k, kv_full, sK, 0, kb * TILE, h, b);
      tma_load_4d(&map_v, kv_full, sV, 0, kb * TILE, h, b);
      for (int i = 0; i < nq; ++i) {
        const int st = i % BW_RING;
        mbar_wait(&r_empty[st], ((i / BW_RING) & 1) ^ 1);
        mbar_expect_tx(&r_full[st], 2 * TILE_BYTES);
        tma_load_4d(&map_q, &r_full[st], sR + (2 * st) * TILE_BYTES, 0, i * TILE, h, b);
        tma_load_4d(&map_do, &r_full[st], sR + (2 * st + 1) * TILE_BYTES, 0, i * TILE, h, b);
      }
    }
  } else if (warp == 5) {
    if (elect_one()) {
      const uint64_t kmaj = make_desc_base(16, 1024);
      const uint64_t mnmaj = make_desc_base(BLOCK_K * 128, 1024);
      const uint32_t sP_u = smem_u32(sP), sS_u = smem_u32(sS);
      const uint64_t dK_k = kmaj + desc_addr(smem_u32(sK));      // K_j as K-major B (N = keys, K = d)
      const uint64_t dV_k = kmaj + desc_addr(smem_u32(sV));      // V_j as K-major B
      const uint64_t dK_mn = mnmaj + desc_addr(smem_u32(sK));    // K_j as MN-major B (N = d, K = keys)
      const uint32_t id_s = idesc_bf16(kpad, false, false);      // [128 q] x [kpad keys], K = d
      const uint32_t id_t = idesc_bf16(HD, true, true);          // P^T dO / dS^T Q : M = keys (MN-major A), N = d
      const uint32_t id_q = idesc_bf16(HD, false, true);         // dS K : A K-major (K = keys), B MN-major
      mbar_wait(kv_full, 0);
      auto issue_sdp = [&](int i) {
        const int st = i % BW_RING;
        const uint32_t sQ_u = smem_u32(sR + (2 * st) * TILE_BYTES);
        mbar_wait(&r_full[st], (i / BW_RING) & 1);
        tc_fence_after();
        const uint64_t dQ_k = kmaj + desc_addr(sQ_u), dO_k = kmaj + desc_addr(sQ_u + TILE_BYTES);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) tc_mma_bf16(tmem + TM_S, dQ_k + 2 * k, dK_k + 2 * k, id_s, k ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) tc_mma_bf16(tmem + TM_DP, dO_k + 2 * k, dV_k + 2 * k, id_s, k ? 1u : 0u);
        tc_commit(sdp_full);
      };
      issue_sdp(0);
      const uint64_t mn_a = make_desc_base(TILE_BYTES, 1024);   // P / dS: 64-key chunks are 128 rows (16 KB) apart
      const uint64_t aP = mn_a + desc_addr(sP_u), aS = mn_a + desc_addr(sS_u);
      for (int i = 0; i < nq; ++i) {
        const int st = i % BW_RING;
        const uint32_t sQ_u = smem_u32(sR + (2 * st) * TILE_BYTES), sO_u = sQ_u + TILE_BYTES;
        mbar_wait(pds_full, i & 1);                               // P_i, dS_i in smem; S / dP TMEM free again
        if (i > 0) mbar_wait(dq_empty, (i - 1) & 1);              // dQ_{i-1} drained from TMEM
        tc_fence_after();
        // dV += P^T dO_i ; dK += dS^T Q_i : A = P / dS read MN-major (M = keys contiguous, K = q rows),
        // chunk c (64 keys) at +c*TILE_BYTES; B = dO_i / Q_i MN-major (N = d, K = q rows)
        const uint64_t bO = mnmaj + desc_addr(sO_u), bQ = mnmaj + desc_addr(sQ_u);
#pragma unroll
        for (int k = 0; k < TILE / 16; ++k) tc_mma_bf16(tmem + TM_DV, aP + k * 128, bO + k * 128, id_t, (i | k) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < TILE / 16; ++k) tc_mma_bf16(tmem + TM_DK, aS + k * 128, bQ + k * 128, id_t, (i | k) ? 1u : 0u);
        // dQ_i = dS K_j : A = dS K-major (K = keys), B = K_j MN-major
        for (int k = 0; k < kpad / 16; ++k) {
          const uint64_t da = kmaj + desc_addr(sS_u + (k >> 2) * TILE_BYTES + (k & 3) * 32);
          tc_mma_bf16(tmem + TM_DQ, da, dK_mn + (uint64_t)(k * 128), id_q, k ? 1u : 0u);
        }
        tc_commit(&r_empty[st]);
        tc_commit(dq_full);
        // next tile's S / dP go out now: their MMAs and the softmax that follows overlap the dQ_i drain.
        // (P / dS smem is only rewritten after dq_full(i), i.e. after the MMAs above have read it.)
        if (i + 1 < nq) issue_sdp(i + 1);
      }
      tc_commit(fin_full);
    }
  } else {
    // ============================ softmax / gradient warps ============================
    const int row = warp * 32 + lane;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t sP_u = smem_u32(sP), sS_u = smem_u32(sS);
    const size_t bh_off = ((size_t)b * p.H + h) * p.S;
    for (int i = 0; i < nq; ++i) {
      const int qrow = i * TILE + row;
      const bool qok = qrow < p.S;
      const float lse2 = qok ? p.lse[bh_off + qrow] * 1.4426950408889634f : 0.f;
      const float dl = qok ? p.delta[bh_off + qrow] : 0.f;
      mbar_wait(sdp_full, i & 1);
      tc_fence_after();
      if (i > 0) {                                       // P / dS smem of tile i-1 must have been consumed
        // (dq_full of tile i-1 was waited below before we got here)
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c * 32 < kpad) {
          uint32_t rs[32], rp[32];
          tc_ld_32x32b_x32(tmem + TM_S + lane_base + c * 32, rs);
          tc_ld_32x32b_x32(tmem + TM_DP + lane_base + c * 32, rp);
          tc_wait_ld();
          float pv[32], ds[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const bool ok = qok && (c * 32 + k < kvalid);
            const float pe = ok ? ex2(fmaf(__uint_as_float(rs[k]), p.scale_log2, -lse2)) : 0.f;
            pv[k] = pe;
            ds[k] = pe * (__uint_as_float(rp[k]) - dl) * p.scale;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int u = (c & 1) * 4 + q;
            const uint32_t off = (c >> 1) * TILE_BYTES + row * 128 + ((u ^ (row & 7)) << 4);
            sts128(sP_u + off, pack8(pv + q * 8));
            sts128(sS_u + off, pack8(ds + q * 8));
          }
        }
      }
      tc_fence_before();
      fence_async_cta();
      mbar_arrive_cnt(pds_full);
      // ---- dQ_i: TMEM -> fp32 RED.ADD (v4) into the workspace
      mbar_wait(dq_full, i & 1);
      tc_fence_after();
      float* dst = p.dq_acc + (size_t)b * p.dq_sb + (size_t)h * p.dq_sh + (size_t)qrow * p.dq_ss;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tc_ld_32x32b_x32(tmem + TM_DQ + lane_base + c * 32, r);
        tc_wait_ld();
        if (qok) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c * 32 + q * 4),
                         "f"(__uint_as_float(r[4 * q])), "f"(__uint_as_float(r[4 * q + 1])),
                         "f"(__uint_as_float(r[4 * q + 2])), "f"(__uint_as_float(r[4 * q + 3]))
                         : "memory");
        }
      }
      tc_fence_before();
      mbar_arrive_cnt(dq_empty);
    }
    // ---- dK_j, dV_j: rows = keys
    mbar_wait(fin_full, 0);
    tc_fence_after();
    const int krow = kb * TILE + row;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      float v[HD];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tc_ld_32x32b_x32(tmem + (which ? TM_DK : TM_DV) + lane_base + c * 32, r);
        tc_wait_ld();
#pragma unroll
        for (int k = 0; k < 32; ++k) v[c * 32 + k] = __uint_as_float(r[k]);
      }
      if (krow < p.S) {
        __nv_bfloat16* base = which ? p.dk + (size_t)b * p.dk_sb + (size_t)h * p.dk_sh + (size_t)krow * p.dk_ss
                                    : p.dv + (size_t)b * p.dv_sb + (size_t)h * p.dv_sh + (size_t)krow * p.dv_ss;
        uint4* dst = reinterpret_cast<uint4*>(base);
#pragma unroll
        for (int q = 0; q < 8; ++q) dst[q] = pack8(v + q * 8);
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
thread_local char g_err[512];

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
  return 0;
}
// {64 d, S, H, B} view with element strides (ss, sh, sb); box {64, 128, 1, 1}
int make_qkv_map(CUtensorMap* m, const void* ptr, int B, int H, int S, long long sb, long long sh, long long ss) {
  if ((ss % 8) || (sh % 8) || (sb % 8) || ((uintptr_t)ptr & 15)) return fail("attention operands must be 16-byte aligned");
  cuuint64_t dims[4] = {(cuuint64_t)HD, (cuuint64_t)S, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)ss * 2, (cuuint64_t)sh * 2, (cuuint64_t)sb * 2};
  cuuint32_t box[4] = {64, TILE, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(attention) failed", (int)r);
  return 0;
}

}  // namespace

extern "C" {

const char* b200dp_attn_last_error() { return g_err; }

// strides: element strides {batch, head, seq} of each tensor (head dim 64 contiguous)
int b200dp_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int S, int D,
                    const long long* qs, const long long* ks, const long long* vs, const long long* os, float scale,
                    unsigned long long stream) {
  if (ensure_init()) return -1;
  if (D != HD) return fail("head dim must be 64");
  CUtensorMap mq, mk, mv;
  if (make_qkv_map(&mq, q, B, H, S, qs[0], qs[1], qs[2]) || make_qkv_map(&mk, k, B, H, S, ks[0], ks[1], ks[2]) ||
      make_qkv_map(&mv, v, B, H, S, vs[0], vs[1], vs[2]))
    return -1;
  if ((os[0] % 8) || (os[1] % 8) || (os[2] % 8) || ((uintptr_t)o & 15)) return fail("output must be 16-byte aligned");
  AttnFwdParams p;
  p.B = B; p.H = H; p.S = S;
  p.q_tiles = (S + TILE - 1) / TILE;
  p.kv_blocks = (S + TILE - 1) / TILE;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.o = reinterpret_cast<__nv_bfloat16*>(o);
  p.o_sb = os[0]; p.o_sh = os[1]; p.o_ss = os[2];
  p.lse = lse;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FW_SMEM);
    if (e != cudaSuccess) return fail(cudaGetErrorString(e), (int)e);
    attr_set = true;
  }
  attn_fwd_kernel<<<B * H * p.q_tiles, AT, FW_SMEM, (cudaStream_t)(uintptr_t)stream>>>(mq, mk, mv, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(cudaGetErrorString(e), (int)e);
  return 0;
}

// dq_acc: fp32 workspace (strides dqs, 64 contiguous) zeroed by the caller; delta: [B][H][S] fp32 workspace
int b200dp_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
                    float* delta, float* dq_acc, void* dk, void* dv, int B, int H, int S, int D, const long long* qs,
                    const long long* ks, const long long* vs, const long long* os, const long long* dos,
                    const long long* dqs, const long long* dks, const long long* dvs, float scale,
                    unsigned long long stream) {
  if (ensure_init()) return -1;
  if (D != HD) return fail("head dim must be 64");
  cudaStream_t st = (cudaStream_t)(uintptr_t)stream;
  CUtensorMap mq, mk, mv, mdo;
  if (make_qkv_map(&mq, q, B, H, S, qs[0], qs[1], qs[2]) || make_qkv_map(&mk, k, B, H, S, ks[0], ks[1], ks[2]) ||
      make_qkv_map(&mv, v, B, H, S, vs[0], vs[1], vs[2]) || make_qkv_map(&mdo, dout, B, H, S, dos[0], dos[1], dos[2]))
    return -1;
  const long long rows = (long long)B * H * S;
  attn_delta_kernel<<<(unsigned)((rows + 31) / 32), 256, 0, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(o), reinterpret_cast<const __nv_bfloat16*>(dout), delta, B, H, S, os[0],
      os[1], os[2], dos[0], dos[1], dos[2]);
  AttnBwdParams p;
  p.B = B; p.H = H; p.S = S;
  p.q_tiles = (S + TILE - 1) / TILE;
  p.kv_blocks = (S + TILE - 1) / TILE;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.lse = lse; p.delta = delta; p.dq_acc = dq_acc;
  p.dq_sb = dqs[0]; p.dq_sh = dqs[1]; p.dq_ss = dqs[2];
  if ((dqs[0] % 4) || (dqs[1] % 4) || (dqs[2] % 4)) return fail("dq workspace strides must be multiples of 4");
  p.dk = reinterpret_cast<__nv_bfloat16*>(dk); p.dv = reinterpret_cast<__nv_bfloat16*>(dv);
  p.dk_sb = dks[0]; p.dk_sh = dks[1]; p.dk_ss = dks[2];
  p.dv_sb = dvs[0]; p.dv_sh = dvs[1]; p.dv_ss = dvs[2];
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BW_SMEM);
    if (e != cudaSuccess) return fail(cudaGetErrorString(e), (int)e);
    attr_set = true;
  }
  attn_bwd_kernel<<<B * H * p.kv_blocks, AT, BW_SMEM, st>>>(mq, mk, mv, mdo, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(cudaGetErrorString(e), (int)e);
  return 0;
}

}  // extern "C"
