// Shared sm_100a device helpers: mbarrier / TMA / tcgen05 PTX wrappers, UMMA descriptors and the
// fused TMEM->register->smem->TMA-store epilogue used by the GEMM (gemm_sm100.cu), the implicit-GEMM
// convolution (conv_sm100.cu) and the attention kernels (attn_sm100.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;          // 64 bf16 = 128 bytes = one 128B swizzle atom
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 320;     // 10 warps: TMA, MMA, 8 epilogue (2 per TMEM lane quarter)
constexpr int EPI_WARPS = 8;

struct GemmParams {
  int M, N, K;
  int ldc;                 // elements
  int num_m_blocks, num_n_blocks, num_k_blocks;
  int splits;              // split-K factor (>=1)
  int act;                 // 0 none, 1 relu, 2 gelu(erf), 3 *gelu'(aux), 4 *(aux>0)  [aux = residual ptr]
  int out_mode;            // 0: bf16 store, 1: fp32 atomic add (split-K / accumulate), 2: fp32 store
  void* C;
  const void* bias;        // bf16 [N] or nullptr
  const void* bias_f32;    // fp32 [N] or nullptr
  const void* residual;    // bf16 [M, ldc] or nullptr (added after act; or `aux` for act 3/4)
  void* preact;            // optional bf16 [M, ldc]: pre-activation values (saved for backward)
  int tma_store;           // out_mode 0: stage the tile in swizzled smem and write it with TMA bulk stores
  int n_fastest;           // tile order: consecutive CTAs walk the N blocks of one M block first (A tile is
                           // fetched from HBM once and re-used from L2 while the whole B matrix stays in L2)
  float alpha;
  const unsigned char* res_mask;   // optional, with a plain residual (act 0..2): bit j of byte [row][col/8] keeps
                           // residual element (row, 8*(col/8)+j) — the ReLU sign bits of the block output whose
                           // skip-branch gradient the residual is (N % 64 == 0 required)
  float* stats;            // optional fp32 [2][N]: per-column sum | sum of squares of the bf16 OUTPUT (the batch
                           // statistics of the BatchNorm that follows), accumulated by the epilogue
};

// Shared-memory accumulators for GemmParams::stats, after the barriers: one PRIVATE region per epilogue warp
// (sum[128] | sumsq[128] floats for the <= 128 tile columns the warp drains).  Every address is only ever
// touched by one lane, so accumulation is plain ld/add/st — fp32 atomicAdd on shared memory compiles to a
// CAS loop (ATOMS.CAST.SPIN) and the four lane-quarter warps of a tile hit the same columns.
constexpr int STATS_MAX_N = 2048;
constexpr int STATS_WARP_FLOATS = 256;
constexpr int STATS_SMEM_BYTES = 8 * STATS_WARP_FLOATS * 4;

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // try_wait suspends for a bounded time per attempt; a %globaltimer watchdog (4 s) turns a protocol
  // bug (lost arrive / wrong phase) into a trap ("unspecified launch failure") instead of a GPU hang.
  uint32_t done = 0;
  unsigned long long t0 = 0;
  for (uint32_t tries = 0; !done; ++tries) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!done && (tries & 1023u) == 1023u) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
// 4D tiled loads/stores (NHWC activations as {C, W, H, N} tensors): out-of-bounds box elements — negative
// or past-the-end coordinates, i.e. the convolution's zero padding — are zero-filled on load and skipped
// on store by the TMA unit itself.
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], "
      "[%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2,
                                             int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld_32x32b_x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 64-bit UMMA shared-memory descriptor (sm_100: version=1), 128B swizzle.
//   K-major  : rows of 128 B (64 bf16 of K), 8-row groups 1024 B apart (SBO); LBO unused (=1).
//   MN-major : [k rows][64 mn elements] atoms of 128 B rows; 8-row k-groups SBO=1024 B apart,
//              64-element mn chunks LBO bytes apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                              uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;   // LayoutType::SWIZZLE_128B
  return d;
}

// One elected lane of a converged warp (PTX elect.sync).  Code guarded by it is known to the compiler
// to run in a single thread, so tcgen05 / TMA operands stay in uniform registers instead of going
// through a per-instruction ELECT + R2UR + BRA.U.ANY sequence (what `lane == 0` compiles to).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// Descriptor without the start address (bits 0..13): OR / add `(smem_addr >> 4)` to place it; advancing
// along K or by whole rows is then a 64-bit add of `(bytes >> 4)` — no rebuild in the MMA issue loop,
// which is a single thread and the critical path of every tcgen05 kernel here (profiles/ncu_conv_*.md).
__device__ __forceinline__ uint64_t make_desc_base(uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return make_desc(0, lbo_bytes, sbo_bytes);
}
__device__ __forceinline__ uint64_t desc_addr(uint32_t smem_addr) { return (uint64_t)((smem_addr & 0x3FFFF) >> 4); }

template <int BN, bool A_MN, bool B_MN>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
  uint32_t d = 0;
  d |= 1u << 4;                      // D format: F32
  d |= 1u << 7;                      // A format: BF16
  d |= 1u << 10;                     // B format: BF16
  d |= (A_MN ? 1u : 0u) << 15;       // A major
  d |= (B_MN ? 1u : 0u) << 16;       // B major
  d |= (uint32_t)(BN >> 3) << 17;    // N
  d |= (uint32_t)(BLOCK_M >> 4) << 24;  // M
  return d;
}

// erf via Abramowitz-Stegun 7.1.26 (|err| < 1.5e-7, far below bf16 resolution): 1 rcp + 1 ex2 + 6 fma
// instead of the ~40-instruction erff — the epilogue, not the MMA, bounds the GELU GEMMs.
__device__ __forceinline__ float fast_erf(float x) {
  const float ax = fabsf(x);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float y = 1.0f - poly * t * __expf(-ax * ax);
  return copysignf(y, x);
}
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + fast_erf(x * 0.70710678118654752f));
}

__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + fast_erf(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

__device__ __forceinline__ void unpack8(const uint4& x, float* f) {
  const uint32_t wv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    f[2 * t] = __uint_as_float(wv[t] << 16);
    f[2 * t + 1] = __uint_as_float(wv[t] & 0xffff0000u);
  }
}

__device__ __forceinline__ uint4 pack8(const float* v) {
  uint32_t wv[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * t], v[2 * t + 1]);
    wv[t] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(wv[0], wv[1], wv[2], wv[3]);
}

template <int BN>
struct Cfg {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BN * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 3 : ((BN == 128) ? 4 : 6);
  static constexpr int TMEM_COLS = 2 * BN;       // double-buffered fp32 accumulator
  static constexpr int STORE_BYTES = EPI_WARPS * 2 * 4096;   // per epilogue warp: out + preact staging (32 rows x 128 B each)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STORE_BYTES + 1024 /*align*/ + 256 /*barriers*/ +
                                    STATS_SMEM_BYTES;
};

// Epilogue of one 32-row slab (this warp's TMEM lane quarter) of a 128 x BN accumulator: TMEM ->
// registers -> alpha/bias/activation/residual -> bf16 via swizzled smem + TMA bulk store, or fp32
// store / atomic add.  Shared by the 1-CTA and the 2-CTA (cta_group::2) kernels.
// Where a 32-row slab goes.  rank4 == 0: rows m_row0.. of the row-major [M, ldc] matrix (2D map).
// rank4 == 1 (convolution): the slab is the {64 c, sw, sh, sn} sub-box at pixel (w, h, n) of an NHWC
// tensor; c_ptr/ld override p.C/p.ldc for the fp32 modes (per-tap column offset of the wgrad output).
struct StoreAt {
  int rank4, w, h, n;
  void* c_ptr;
  // rank4 only (statistics of a convolution output): the slab is an {sw, sh, .} pixel box of which only
  // vw x vh x vn pixels lie inside the image / batch — rows outside are computed from partly valid taps
  // (NOT zero) and must not be counted; the TMA store clips them by itself.
  int sw, sh, vw, vh, vn;
};

// explicit shared-space accesses (a generic pointer into shared memory costs 64-bit address arithmetic
// and the generic LD/ST path: measured, the epilogue is issue-bound)
__device__ __forceinline__ void epi_sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 epi_lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
// prmt.b32 in its default mode: selector nibble bit 3 = replicate the selected byte's sign bit
__device__ __forceinline__ uint32_t prmt_sign(uint32_t a, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(0u), "r"(sel));
  return d;
}
__device__ __forceinline__ uint4 pack8r(const uint32_t* r) {   // 8 fp32 bit patterns -> 8 bf16
  uint4 o;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o.x) : "f"(__uint_as_float(r[1])), "f"(__uint_as_float(r[0])));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o.y) : "f"(__uint_as_float(r[3])), "f"(__uint_as_float(r[2])));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o.z) : "f"(__uint_as_float(r[5])), "f"(__uint_as_float(r[4])));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o.w) : "f"(__uint_as_float(r[7])), "f"(__uint_as_float(r[6])));
  return o;
}

template <int BN>
__device__ __forceinline__ void epilogue_rows(const GemmParams& p, const CUtensorMap* map_c,
                                              const CUtensorMap* map_z, uint32_t tmem_base, int acc, int q,
                                              int lane, int m_row0, int n_idx, int c_begin, int c_end,
                                              uint8_t* my_store, const StoreAt at = StoreAt{0, 0, 0, 0, nullptr, 32, 1, 32, 1, 1},
                                              float* s_stats = nullptr) {
  // A 64-column chunk (one 128-byte bf16 row per lane, one TMA store box) is produced in two 32-column
  // halves: 32 accumulator values + 32 results live per thread instead of 64 + 64 — the previous
  // single-pass version spilled ~300 B per thread at the 168-register budget of a 320-thread CTA, and the
  // epilogue, not the MMA, bounds every K <= 512 GEMM / convolution here.
  const int row = m_row0 + lane;
  const bool row_ok = row < p.M;
  const bool to_tma = p.out_mode == 0 && p.tma_store;
  const bool z_tma = p.preact != nullptr && to_tma;
  const bool res_smem = p.residual != nullptr && p.preact == nullptr && p.out_mode != 1;
  // the common case (plain bf16 output, optionally with BN statistics): no per-element work at all
  const bool plain = to_tma && p.alpha == 1.0f && p.bias == nullptr && p.bias_f32 == nullptr &&
                     p.preact == nullptr && p.act == 0 && p.residual == nullptr;
  const bool res_only = to_tma && res_smem && p.alpha == 1.0f && p.bias == nullptr && p.bias_f32 == nullptr &&
                        p.act == 0;
  const uint32_t store_s = smem_u32(my_store);
  const uint32_t lane_row = (uint32_t)lane * 128u;
  const uint32_t lsw = (uint32_t)(lane & 7);
#pragma unroll 1
  for (int c0 = c_begin; c0 < c_end; c0 += 64) {
    const int col0 = n_idx + c0;
    if (col0 >= p.N) continue;                       // warp-uniform
    const int ncols = min(64, p.N - col0);           // N % 8 == 0 is enforced by the host
    // Residual / auxiliary operand of the slab: COALESCED 16-byte loads (a warp instruction covers 4 rows
    // x 128 B), issued before the TMEM loads so that their latency overlaps, then transposed to
    // row-per-lane through the second (pre-activation) staging buffer.
    uint4 resv[8];
    if (res_smem) {
      const __nv_bfloat16* rbase = reinterpret_cast<const __nv_bfloat16*>(p.residual);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = i * 4 + (lane >> 3), u = lane & 7;
        resv[i] = (m_row0 + rr < p.M && u * 8 < ncols)
                      ? *reinterpret_cast<const uint4*>(rbase + (size_t)(m_row0 + rr) * p.ldc + col0 + u * 8)
                      : make_uint4(0, 0, 0, 0);
      }
      if (p.res_mask != nullptr) {
        // ReLU sign bits of the residual (1 byte per 8 channels): bit j -> 16-bit lane j.  Byte k of
        // ((bits * 0x01010101) & 0x08040201) + 0x7f7f7f7f has its MSB set iff bit k is, and PRMT's
        // sign-replicate mode turns that MSB into 0x00 / 0xff bytes.
        uint32_t mbits[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + (lane >> 3), u = lane & 7;
          mbits[i] = (m_row0 + rr < p.M && u * 8 < ncols)
                         ? (uint32_t)p.res_mask[(size_t)(m_row0 + rr) * (size_t)(p.N >> 3) + (size_t)((col0 >> 3) + u)]
                         : 0u;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t rep = mbits[i] * 0x01010101u;
          const uint32_t lo = (rep & 0x08040201u) + 0x7f7f7f7fu, hi = (rep & 0x80402010u) + 0x7f7f7f7fu;
          resv[i].x &= prmt_sign(lo, 0x9988u);
          resv[i].y &= prmt_sign(lo, 0xbbaau);
          resv[i].z &= prmt_sign(hi, 0x9988u);
          resv[i].w &= prmt_sign(hi, 0xbbaau);
        }
      }
    }
    // Staging: two 4 KB buffers per warp.  When the second one is not needed for the pre-activation tile or
    // the residual transpose, consecutive chunks ALTERNATE between them and only wait for the store issued
    // two chunks ago (cp.async.bulk.wait_group.read 1) — otherwise every chunk stalls on its predecessor's
    // bulk store reading shared memory.
    const bool dbuf = to_tma && !z_tma && !res_smem;
    // buffer parity must alternate over the sequence of chunks THIS warp stages, across tiles: with an even
    // number of chunks per tile the chunk index does it, with one chunk per tile the (alternating) accumulator
    // index does
    const int par = (((c0 - c_begin) >> 6) + acc * (((c_end - c_begin) >> 6) & 1)) & 1;
    const uint32_t out_s = store_s + ((dbuf && par) ? 4096u : 0u);
    if (res_smem) {
      if (lane == 0) tma_store_wait_read<0>();       // (second buffer is about to be rewritten)
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = i * 4 + (lane >> 3), u = lane & 7;
        epi_sts128(store_s + 4096u + rr * 128 + ((u ^ (rr & 7)) << 4), resv[i]);
      }
      __syncwarp();
    }
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0);
    if (plain) {
      // ---- fast path: TMEM -> bf16 -> swizzled staging rows, nothing else ----
      uint32_t r[32];
      tc_ld_32x32b_x32(taddr, r);
      if (lane == 0) {       // the bulk store that last read this buffer must be done reading it
        if (dbuf) tma_store_wait_read<1>();
        else tma_store_wait_read<0>();
      }
      tc_wait_ld();
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j) epi_sts128(out_s + lane_row + ((j ^ lsw) << 4), pack8r(r + j * 8));
      if (ncols > 32) {
        tc_ld_32x32b_x32(taddr + 32, r);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 4; ++j) epi_sts128(out_s + lane_row + (((4 + j) ^ lsw) << 4), pack8r(r + j * 8));
      }
    } else if (res_only) {
      // ---- skip-gradient path of the dgrad GEMMs: bf16(acc + residual), packed fp32x2 adds ----
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        if (half * 32 >= ncols) break;
        uint32_t r[32];
        tc_ld_32x32b_x32(taddr + half * 32, r);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t off = lane_row + (((half * 4 + j) ^ lsw) << 4);
          const uint4 rv = epi_lds128(store_s + 4096u + off);
          const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
          uint4 o;
          uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            unsigned long long a2, b2;
            asm("mov.b64 %0, {%1, %2};" : "=l"(a2) : "r"(r[j * 8 + 2 * w]), "r"(r[j * 8 + 2 * w + 1]));
            asm("mov.b64 %0, {%1, %2};" : "=l"(b2) : "r"(rw[w] << 16), "r"(rw[w] & 0xffff0000u));
            asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a2) : "l"(b2));
            uint32_t s0, s1;
            asm("mov.b64 {%0, %1}, %2;" : "=r"(s0), "=r"(s1) : "l"(a2));
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(ow[w]) : "f"(__uint_as_float(s1)), "f"(__uint_as_float(s0)));
          }
          epi_sts128(out_s + off, o);      // the store that read this buffer was waited for above (res_smem)
        }
      }
    } else {
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      const int hc = half * 32;                      // first column of this half inside the chunk
      if (hc >= ncols) break;                        // warp-uniform
      const int hcols = min(32, ncols - hc);
      uint32_t r[32];
      tc_ld_32x32b_x32(taddr + hc, r);
      tc_wait_ld();
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
      if (p.alpha != 1.0f) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= p.alpha;
      }
      if (p.out_mode != 1) {
        if (p.bias != nullptr) {
          const __nv_bfloat16* b = reinterpret_cast<const __nv_bfloat16*>(p.bias) + col0 + hc;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < hcols) v[i] += __bfloat162float(b[i]);
        } else if (p.bias_f32 != nullptr) {
          const float* b = reinterpret_cast<const float*>(p.bias_f32) + col0 + hc;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < hcols) v[i] += b[i];
        }
        if (p.preact != nullptr) {
          if (z_tma) {            // pre-activation tile -> second staging buffer (stored by TMA with the output)
            if (half == 0) {
              if (lane == 0) tma_store_wait_read<0>();
              __syncwarp();
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
              epi_sts128(store_s + 4096u + lane_row + (((half * 4 + j) ^ lsw) << 4), pack8(v + j * 8));
          } else if (row_ok) {
            uint4* pp = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.preact) +
                                                 (size_t)row * p.ldc + col0 + hc);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (j * 8 < hcols) pp[j] = pack8(v + j * 8);
          }
        }
        if (p.act == 1) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
        } else if (p.act == 2) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
        }
        if (p.residual != nullptr && (row_ok || res_smem)) {
          const uint4* rp = reinterpret_cast<const uint4*>(
              reinterpret_cast<const __nv_bfloat16*>(p.residual) + (size_t)row * p.ldc + col0 + hc);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j * 8 < hcols) {
              float a[8];
              unpack8(res_smem ? epi_lds128(store_s + 4096u + lane_row + (((half * 4 + j) ^ lsw) << 4)) : rp[j], a);
#pragma unroll
              for (int t = 0; t < 8; ++t) {
                if (p.act == 3) v[j * 8 + t] *= gelu_erf_grad(a[t]);
                else if (p.act == 4) v[j * 8 + t] = a[t] > 0.0f ? v[j * 8 + t] : 0.0f;
                else v[j * 8 + t] += a[t];
              }
            }
          }
        }
      }
      if (to_tma) {
        if (half == 0) {       // the bulk store that last read this buffer must be done reading it
          if (lane == 0) {
            if (dbuf) tma_store_wait_read<1>();
            else tma_store_wait_read<0>();
          }
          __syncwarp();
        }
        // this half of the 32 x 128 B swizzled staging rows (conflict-free 16-byte stores)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          epi_sts128(out_s + lane_row + (((half * 4 + j) ^ lsw) << 4), pack8(v + j * 8));
      } else if (p.out_mode == 1) {
        // split-K accumulation: transpose the 32 x 32 fp32 half through the staging buffer so that a
        // warp-level RED covers four 128-byte row segments with 16-byte vectors (red.global.add.v4.f32)
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          epi_sts128(store_s + lane_row + ((j ^ lsw) << 4),
                 make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
                            __float_as_uint(v[4 * j + 3])));
        __syncwarp();
        float* cbase = reinterpret_cast<float*>(at.c_ptr ? at.c_ptr : p.C);
        const int ch = lane & 7;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rr = it * 4 + (lane >> 3);
          const uint4 t = epi_lds128(store_s + rr * 128 + ((ch ^ (rr & 7)) << 4));
          if (m_row0 + rr < p.M && ch * 4 < hcols) {
            float* dst = cbase + (size_t)(m_row0 + rr) * p.ldc + col0 + hc + ch * 4;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(t.x), "r"(t.y), "r"(t.z),
                         "r"(t.w)
                         : "memory");
          }
        }
        __syncwarp();
      } else if (row_ok) {
        if (p.out_mode == 0) {
          uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.C) + (size_t)row * p.ldc + col0 + hc);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j * 8 < hcols) dst[j] = pack8(v + j * 8);
        } else {   // out_mode 2: fp32 store
          float* dst = reinterpret_cast<float*>(at.c_ptr ? at.c_ptr : p.C) + (size_t)row * p.ldc + col0 + hc;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (j * 4 < hcols)
              reinterpret_cast<float4*>(dst)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
    }
    }
    if (to_tma) {
      fence_async_smem();
      __syncwarp();
      if (s_stats != nullptr) {
        // batch statistics of the following BatchNorm: column sums of the bf16 values just staged.  Lane l
        // owns columns 2l, 2l+1 of this 64-column chunk and walks the 32 staged rows (one conflict-free
        // 4-byte shared load per row, packed fp32x2 add / fma); partials go to the warp's private shared
        // accumulators.  GEMM rows past M come from zero-filled operands and add nothing; rows of a
        // convolution tile outside the image are masked (they see partly valid taps).
        unsigned long long sum2 = 0ull, sq2 = 0ull;
        const uint32_t base = out_s + (uint32_t)((lane & 3) << 2);
        const uint32_t u = (uint32_t)(lane >> 2);
        if (at.rank4) {      // lane r decides for slab row r
          const int rw = lane % at.sw, rh = (lane / at.sw) % at.sh, rn = lane / (at.sw * at.sh);
          const uint32_t rows_ok = __ballot_sync(0xffffffffu, rw < at.vw && rh < at.vh && rn < at.vn);
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) {
            uint32_t w;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(base + rr * 128 + ((u ^ (rr & 7)) << 4)));
            if (!((rows_ok >> rr) & 1u)) w = 0u;
            unsigned long long ab;
            asm("mov.b64 %0, {%1, %2};" : "=l"(ab) : "r"(w << 16), "r"(w & 0xffff0000u));
            asm("add.rn.f32x2 %0, %0, %1;" : "+l"(sum2) : "l"(ab));
            asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(sq2) : "l"(ab));
          }
        } else {
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) {
            uint32_t w;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(base + rr * 128 + ((u ^ (rr & 7)) << 4)));
            unsigned long long ab;
            asm("mov.b64 %0, {%1, %2};" : "=l"(ab) : "r"(w << 16), "r"(w & 0xffff0000u));
            asm("add.rn.f32x2 %0, %0, %1;" : "+l"(sum2) : "l"(ab));
            asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(sq2) : "l"(ab));
          }
        }
        // lane-private slots of the warp's region (columns past N hold zeros: zero-filled B rows)
        const uint32_t ps = smem_u32(s_stats) + (uint32_t)(((c0 - c_begin) + 2 * lane) << 2);
        unsigned long long a0, a1;
        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(a0) : "r"(ps));
        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(a1) : "r"(ps + (STATS_WARP_FLOATS / 2) * 4));
        asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a0) : "l"(sum2));
        asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a1) : "l"(sq2));
        asm volatile("st.shared.b64 [%0], %1;" ::"r"(ps), "l"(a0) : "memory");
        asm volatile("st.shared.b64 [%0], %1;" ::"r"(ps + (STATS_WARP_FLOATS / 2) * 4), "l"(a1) : "memory");
      }
      if (lane == 0) {
        const uint8_t* buf = my_store + (out_s - store_s);
        if (at.rank4) {
          tma_store_4d(map_c, buf, col0, at.w, at.h, at.n);
        } else {
          tma_store_2d(map_c, buf, col0, m_row0);
          if (p.preact != nullptr) tma_store_2d(map_z, buf + 4096, col0, m_row0);
        }
        tma_store_commit();
      }
    }
  }
}

// shared accumulators of GemmParams::stats: zero (all threads, before the kernel's first __syncthreads) and
// flush (all 8 epilogue warps together — named barrier 2 among the 256 of them — whenever the CTA moves to
// another column block and after its last tile): the four lane-quarter regions of a column are summed and
// added to global memory once.
__device__ __forceinline__ void stats_zero(float* s_stats, int nthreads) {
  for (int i = threadIdx.x; i < 8 * STATS_WARP_FLOATS; i += nthreads) s_stats[i] = 0.f;
}
template <int BN>
__device__ __forceinline__ void stats_flush(const GemmParams& p, float* s_stats, int n_idx, int epi_tid) {
  constexpr int HALF = (BN >= 128) ? BN / 2 : BN;         // columns per epilogue warp
  asm volatile("bar.sync 2, 256;" ::: "memory");
  if (epi_tid < BN) {
    const int h = epi_tid / HALF, l = epi_tid % HALF;
    const float* r = s_stats + (h * 4) * STATS_WARP_FLOATS + l;
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s += r[k * STATS_WARP_FLOATS];
      q += r[k * STATS_WARP_FLOATS + STATS_WARP_FLOATS / 2];
    }
    if (n_idx + epi_tid < p.N) {
      atomicAdd(p.stats + n_idx + epi_tid, s);
      atomicAdd(p.stats + p.N + n_idx + epi_tid, q);
    }
  }
  asm volatile("bar.sync 2, 256;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) s_stats[i * 256 + epi_tid] = 0.f;
  asm volatile("bar.sync 2, 256;" ::: "memory");
}

// cuTensorMapEncode* is a driver-API call: it fails with CUDA_ERROR_INVALID_CONTEXT on a thread that
// has not yet bound the primary context (e.g. an autograd worker whose first GPU op is one of ours).
inline void bind_primary_context() {
  static thread_local bool bound = false;
  if (!bound) {
    cudaFree(nullptr);
    bound = true;
  }
}

}  // namespace
