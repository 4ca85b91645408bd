// K6 — fused linear head of the reference model:
//
//     last = lstm_out[:, T-1, :]                 (the reference's index_select, folded into the load)
//     pred = W3 (W2 (W1 last + b1) + b2) + b3     (no activations: app/torch_train.py:199-205)
//
// fp32 (the reference runs fp32), batch 32, 256 -> 256 -> 64 -> 1: 2.6 MFLOP — a latency problem,
// not a throughput one.  cuBLAS needs three GEMM launches + three bias adds forward and six GEMMs +
// three reductions backward; here the forward is ONE kernel (one CTA per sample, intermediates in
// shared memory) and the backward is TWO (per-sample chain, then all weight/bias gradients).
// Warps own output features and stride the reduction dimension across lanes, so every weight row
// is read with coalesced 128-byte accesses and reduced with shuffles.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace {

constexpr int HT = 256;   // threads per block (8 warps)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// out[j] = bias[j] + sum_k in[k] * W[j][k]   for j in [0, N)  (W row-major [N][K]); in/out in smem.
__device__ __forceinline__ void dense_rows(const float* __restrict__ W, const float* __restrict__ bias,
                                           const float* in, float* out, int N, int K) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = HT / 32;
  for (int j = warp; j < N; j += nw) {
    const float* w = W + (size_t)j * K;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(in[k], __ldg(w + k), acc);
    acc = warp_sum(acc);
    if (lane == 0) out[j] = acc + (bias ? __ldg(bias + j) : 0.f);
  }
}

// out[k] = sum_j in[j] * W[j][k]   (transposed product: dX = dY @ W); threads own k (coalesced in k).
__device__ __forceinline__ void dense_cols(const float* __restrict__ W, const float* in, float* out, int N,
                                           int K) {
  for (int k = threadIdx.x; k < K; k += HT) {
    float acc = 0.f;
    for (int j = 0; j < N; ++j) acc = fmaf(in[j], __ldg(W + (size_t)j * K + k), acc);
    out[k] = acc;
  }
}

struct HeadDims {
  int B, K0, N1, N2, N3;   // K0 = hidden size (input features), N1/N2/N3 = layer widths
};

// one CTA per sample
__global__ void __launch_bounds__(HT) head_fwd_kernel(const float* __restrict__ x, long long x_stride,
                                                      const float* W1, const float* b1, const float* W2,
                                                      const float* b2, const float* W3, const float* b3,
                                                      float* a1_out, float* a2_out, float* pred, HeadDims d) {
  extern __shared__ float sm[];
  float* s0 = sm;                 // [K0]
  float* s1 = s0 + d.K0;          // [N1]
  float* s2 = s1 + d.N1;          // [N2]
  float* s3 = s2 + d.N2;          // [N3]
  const int b = blockIdx.x;
  for (int k = threadIdx.x; k < d.K0; k += HT) s0[k] = x[(size_t)b * x_stride + k];
  __syncthreads();
  dense_rows(W1, b1, s0, s1, d.N1, d.K0);
  __syncthreads();
  dense_rows(W2, b2, s1, s2, d.N2, d.N1);
  __syncthreads();
  dense_rows(W3, b3, s2, s3, d.N3, d.N2);
  __syncthreads();
  for (int j = threadIdx.x; j < d.N1; j += HT) a1_out[(size_t)b * d.N1 + j] = s1[j];
  for (int j = threadIdx.x; j < d.N2; j += HT) a2_out[(size_t)b * d.N2 + j] = s2[j];
  for (int j = threadIdx.x; j < d.N3; j += HT) pred[(size_t)b * d.N3 + j] = s3[j];
}

// backward, per sample: da2 = dp @ W3, da1 = da2 @ W2, dx = da1 @ W1 (written at the last timestep)
__global__ void __launch_bounds__(HT) head_bwd_rows_kernel(const float* __restrict__ dpred, const float* W1,
                                                           const float* W2, const float* W3, float* da1_out,
                                                           float* da2_out, float* dx, long long dx_stride,
                                                           HeadDims d) {
  extern __shared__ float sm[];
  float* g3 = sm;                 // [N3]
  float* g2 = g3 + d.N3;          // [N2]
  float* g1 = g2 + d.N2;          // [N1]
  float* g0 = g1 + d.N1;          // [K0]
  const int b = blockIdx.x;
  for (int j = threadIdx.x; j < d.N3; j += HT) g3[j] = dpred[(size_t)b * d.N3 + j];
  __syncthreads();
  dense_cols(W3, g3, g2, d.N3, d.N2);
  __syncthreads();
  dense_cols(W2, g2, g1, d.N2, d.N1);
  __syncthreads();
  dense_cols(W1, g1, g0, d.N1, d.K0);
  __syncthreads();
  for (int j = threadIdx.x; j < d.N2; j += HT) da2_out[(size_t)b * d.N2 + j] = g2[j];
  for (int j = threadIdx.x; j < d.N1; j += HT) da1_out[(size_t)b * d.N1 + j] = g1[j];
  for (int k = threadIdx.x; k < d.K0; k += HT) dx[(size_t)b * dx_stride + k] = g0[k];
}

// backward, weights: dW[j][k] = sum_b dY[b][j] * X[b][k], db[j] = sum_b dY[b][j]; one block per row j of
// the concatenated (layer1 | layer2 | layer3) output features.
__global__ void __launch_bounds__(HT) head_bwd_weights_kernel(
    const float* __restrict__ x, long long x_stride, const float* __restrict__ a1,
    const float* __restrict__ a2, const float* __restrict__ da1, const float* __restrict__ da2,
    const float* __restrict__ dpred, float* dW1, float* db1, float* dW2, float* db2, float* dW3, float* db3,
    HeadDims d) {
  __shared__ float dy[1024];      // dY[:, j] for this output feature (B <= 1024)
  int j = blockIdx.x;
  const float* dY;
  const float* X;
  long long xs;
  float* dW;
  float* db;
  int N, K;
  if (j < d.N1) { dY = da1; N = d.N1; X = x; xs = x_stride; K = d.K0; dW = dW1; db = db1; }
  else if (j < d.N1 + d.N2) { j -= d.N1; dY = da2; N = d.N2; X = a1; xs = d.N1; K = d.N1; dW = dW2; db = db2; }
  else { j -= d.N1 + d.N2; dY = dpred; N = d.N3; X = a2; xs = d.N2; K = d.N2; dW = dW3; db = db3; }
  for (int b = threadIdx.x; b < d.B; b += HT) dy[b] = dY[(size_t)b * N + j];
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += HT) {
    float acc = 0.f;
    for (int b = 0; b < d.B; ++b) acc = fmaf(dy[b], X[(size_t)b * xs + k], acc);
    dW[(size_t)j * K + k] = acc;
  }
  if (threadIdx.x < 32) {
    float s = 0.f;
    for (int b = threadIdx.x; b < d.B; b += 32) s += dy[b];
    s = warp_sum(s);
    if (threadIdx.x == 0) db[j] = s;
  }
}

thread_local char g_err[256];
int fail(const char* what, cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return -1;
}

}  // namespace

extern "C" {

const char* b200dp_lstm_last_error() { return g_err; }

// x: fp32, sample b at x + b*x_stride (so the last LSTM timestep is selected by pointer arithmetic).
int b200dp_head_fwd(const float* x, long long x_stride, const float* W1, const float* b1, const float* W2,
                    const float* b2, const float* W3, const float* b3, float* a1, float* a2, float* pred,
                    int B, int K0, int N1, int N2, int N3, unsigned long long stream) {
  HeadDims d{B, K0, N1, N2, N3};
  const size_t smem = sizeof(float) * (size_t)(K0 + N1 + N2 + N3);
  if (smem > 48 * 1024 || B > 1024) {
    snprintf(g_err, sizeof(g_err), "head dims too large");
    return -1;
  }
  head_fwd_kernel<<<B, HT, smem, (cudaStream_t)(uintptr_t)stream>>>(x, x_stride, W1, b1, W2, b2, W3, b3, a1,
                                                                   a2, pred, d);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("head_fwd launch", e);
  return 0;
}

int b200dp_head_bwd(const float* dpred, const float* x, long long x_stride, const float* a1, const float* a2,
                    const float* W1, const float* W2, const float* W3, float* da1, float* da2, float* dx,
                    long long dx_stride, float* dW1, float* db1, float* dW2, float* db2, float* dW3,
                    float* db3, int B, int K0, int N1, int N2, int N3, unsigned long long stream) {
  HeadDims d{B, K0, N1, N2, N3};
  cudaStream_t st = (cudaStream_t)(uintptr_t)stream;
  const size_t smem = sizeof(float) * (size_t)(K0 + N1 + N2 + N3);
  if (smem > 48 * 1024 || B > 1024) {
    snprintf(g_err, sizeof(g_err), "head dims too large");
    return -1;
  }
  head_bwd_rows_kernel<<<B, HT, smem, st>>>(dpred, W1, W2, W3, da1, da2, dx, dx_stride, d);
  head_bwd_weights_kernel<<<N1 + N2 + N3, HT, 0, st>>>(x, x_stride, a1, a2, da1, da2, dpred, dW1, db1, dW2,
                                                      db2, dW3, db3, d);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("head_bwd launch", e);
  return 0;
}

}  // extern "C"
