// Data-parallel log-likelihood + gradient for ONE parameter vector over millions of rows (BASELINE config 5:
// MLP 16-64-64-1, fp32, 8M rows sharded over the GPUs of a box; the caller all-reduces the partial sums).
//
// Replaces, for large data sets, the same reference path as mlp_static.cuh:
//   eeyore/models/mlp.py:45-50, eeyore/stats/loss.py:1-11, eeyore/models/bayesian_model.py:30-35,
//   eeyore/models/log_target_model.py:15-23 (torch.autograd over [N, 64] activations, ~10 such buffers in HBM)
// and the leapfrog / accept arithmetic of eeyore/samplers/hmc.py:100-170 for a replicated chain state.
//
// Kernel design (compute-bound: 29,056 FLOP vs 68 bytes per row): persistent CTAs, tiles of 128 rows; x / y tiles
// arrive by TMA 1-D bulk copies (double buffered, mbarrier completion); weights live in shared memory; the five
// GEMM-shaped contractions of forward + backward run as register-blocked FP32 FMA loops over shared-memory operands
// stored feature-major so every operand load is a 128-bit LDS; activations never touch HBM.  Weight-gradient tiles are
// accumulated in FP32 registers per tile and folded into FP64 registers, then reduced across CTAs in a fixed order
// (deterministic two-stage reduction).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string>
#include <string.h>
#include "common.cuh"
#include "philox.cuh"
#include "datapar.cuh"
#include "datapar_post.cuh"
#include "../../include/eeyore_b200.h"

#ifndef DP_UNROLL_KMAJOR
#define DP_UNROLL_KMAJOR 16
#endif
#ifndef DP_UNROLL_KMINOR
#define DP_UNROLL_KMINOR 8
#endif

namespace eb {

constexpr int DP_RS = DP_R + 4;      // row stride of feature-major buffers (floats); RS % 32 == 4
constexpr double kLogSqrt2PiD = kDpLogSqrt2Pi;

struct DpSmem {
  alignas(16) unsigned long long bar[2];
  alignas(16) float xs[2][DP_R * DP_D0];   // raw x tiles (row-major), TMA destinations
  alignas(16) float ys[2][DP_R];
  alignas(16) float w0t[DP_D0 * DP_H];     // [j][perm(o)]
  alignas(16) float w1t[DP_H * DP_H];      // [i][perm(o)]   (forward:  k = input unit)
  alignas(16) float w1[DP_H * DP_H];       // [o][perm(i)]   (backward: k = output unit)
  alignas(16) float b0[DP_H], b1[DP_H], w2[DP_H];
  alignas(16) float xt[DP_D0 * DP_RS];     // x tile, feature-major [j][r]
  alignas(16) float A[DP_H * DP_RS];       // H1 -> Delta1, feature-major [f][r]
  alignas(16) float B[DP_H * DP_RS];       // H2 -> Delta2
  alignas(16) float d3[DP_R];              // delta at the head: y - p (0 for padding rows)
  double red[DP_THREADS];
  float b2;
};

// logical feature owned by (group g in 0..15, slot c in 0..3) = g + 16 c; stored at position 4 g + c
__host__ __device__ inline int dp_perm(int f) { return 4 * (f & 15) + (f >> 4); }

__device__ __forceinline__ uint32_t dp_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dp_mbar_init(unsigned long long* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dp_smem_u32(bar)));
}
__device__ __forceinline__ void dp_issue_tile(DpSmem& s, int stage, const float* x, const float* y, long row0, int rows) {
  // bulk copies move multiples of 16 bytes: a ragged tail of y (< 4 values) is read with plain loads by the head step
  const uint32_t xb = (uint32_t)rows * DP_D0 * 4, yb = (uint32_t)(rows & ~3) * 4;
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dp_smem_u32(&s.bar[stage])), "r"(xb + yb) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   dp_smem_u32(s.xs[stage])), "l"(x + row0 * DP_D0), "r"(xb), "r"(dp_smem_u32(&s.bar[stage])) : "memory");
  if (yb)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     dp_smem_u32(s.ys[stage])), "l"(y + row0), "r"(yb), "r"(dp_smem_u32(&s.bar[stage])) : "memory");
}
__device__ __forceinline__ void dp_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "DPWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra DPWAIT_%=;\n\t}" ::"r"(dp_smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ float dp_sigmoid(float z) { return __frcp_rn(1.0f + __expf(-z)); }

// acc[m][n] += sum_k At[k][row0 + m] * Bm[k][col0 + n]    (both operands contiguous along the output index)
template <int TM, int TN, int K>
__device__ __forceinline__ void gemm_kmajor(const float* __restrict__ At, int lda, const float* __restrict__ Bm, int ldb,
                                            int row0, int col0, float (&acc)[TM][TN]) {
  constexpr int kUnroll = DP_UNROLL_KMAJOR;
#pragma unroll kUnroll
  for (int k = 0; k < K; ++k) {
    float a[TM], b[TN];
#pragma unroll
    for (int m = 0; m < TM; m += 4) *reinterpret_cast<float4*>(&a[m]) = *reinterpret_cast<const float4*>(&At[k * lda + row0 + m]);
#pragma unroll
    for (int n = 0; n < TN; n += 4) *reinterpret_cast<float4*>(&b[n]) = *reinterpret_cast<const float4*>(&Bm[k * ldb + col0 + n]);
#pragma unroll
    for (int m = 0; m < TM; ++m)
#pragma unroll
      for (int n = 0; n < TN; ++n) acc[m][n] = fmaf(a[m], b[n], acc[m][n]);
  }
}

// acc[m][n] += sum_{r < R} U[(u0 + 16 m)][r] * V[(v0 + VS n)][r]   (both operands contiguous along r)
template <int TM, int TN, int VS>
__device__ __forceinline__ void gemm_kminor(const float* __restrict__ U, const float* __restrict__ V, int u0, int v0,
                                            float (&acc)[TM][TN]) {
  constexpr int kUnroll = DP_UNROLL_KMINOR;
#pragma unroll kUnroll
  for (int r = 0; r < DP_R; r += 4) {
    float4 a[TM], b[TN];
#pragma unroll
    for (int m = 0; m < TM; ++m) a[m] = *reinterpret_cast<const float4*>(&U[(u0 + 16 * m) * DP_RS + r]);
#pragma unroll
    for (int n = 0; n < TN; ++n) b[n] = *reinterpret_cast<const float4*>(&V[(v0 + VS * n) * DP_RS + r]);
#pragma unroll
    for (int m = 0; m < TM; ++m)
#pragma unroll
      for (int n = 0; n < TN; ++n) {
        acc[m][n] = fmaf(a[m].x, b[n].x, acc[m][n]);
        acc[m][n] = fmaf(a[m].y, b[n].y, acc[m][n]);
        acc[m][n] = fmaf(a[m].z, b[n].z, acc[m][n]);
        acc[m][n] = fmaf(a[m].w, b[n].w, acc[m][n]);
      }
  }
}

// partials: [gridDim.x][DP_P + 1] doubles: [0] = log-likelihood, [1 + j] = d loglik / d theta_j
__global__ void __launch_bounds__(DP_THREADS, 1)
dp_eval_kernel(const float* __restrict__ theta, const float* __restrict__ x, const float* __restrict__ y, long n_rows,
               double* __restrict__ partials) {
  extern __shared__ __align__(16) unsigned char dp_raw[];
  DpSmem& s = *reinterpret_cast<DpSmem*>(dp_raw);
  const int tid = threadIdx.x;
  const long n_tiles = (n_rows + DP_R - 1) / DP_R;

  // ---- stage the weights (permuted so that every thread's operands are contiguous) --------------------------------
  for (int e = tid; e < DP_H * DP_D0; e += DP_THREADS) {  // W0[o][j]
    const int o = e / DP_D0, j = e % DP_D0;
    s.w0t[j * DP_H + dp_perm(o)] = theta[e];
  }
  for (int e = tid; e < DP_H * DP_H; e += DP_THREADS) {   // W1[o][i]
    const int o = e / DP_H, i = e % DP_H;
    const float w = theta[DP_OFF_W1 + e];
    s.w1t[i * DP_H + dp_perm(o)] = w;
    s.w1[o * DP_H + dp_perm(i)] = w;
  }
  if (tid < DP_H) {
    s.b0[tid] = theta[DP_OFF_B0 + tid];
    s.b1[tid] = theta[DP_OFF_B1 + tid];
    s.w2[tid] = theta[DP_OFF_W2 + tid];
  }
  if (tid == 0) {
    s.b2 = theta[DP_OFF_B2];
    dp_mbar_init(&s.bar[0]);
    dp_mbar_init(&s.bar[1]);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // thread roles
  const int ty = tid >> 4, tx = tid & 15;      // k-major GEMMs: rows ty*8 .. +7, features tx + 16 c (stored at 4 tx + c)
  const int ef = tid >> 2, es = tid & 3;       // element-wise passes: feature ef, rows es*32 .. +31
  // persistent FP64 accumulators
  double g1[4][4], g0[2][2];                   // dW1[o = ty + 16 m][i = tx + 16 n], dW0[o = (tid>>3) + 32 m][j = (tid&7) + 8 n]
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int n = 0; n < 4; ++n) g1[m][n] = 0.0;
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < 2; ++n) g0[m][n] = 0.0;
  double gb1 = 0.0, gb0 = 0.0, gw2 = 0.0, gb2 = 0.0, ll = 0.0;
  const int o5 = tid >> 3, j5 = tid & 7;

  long tile = blockIdx.x;
  int stage = 0;
  uint32_t phase[2] = {0u, 0u};
  if (tile < n_tiles && tid == 0) {
    const long r0 = tile * DP_R;
    dp_issue_tile(s, 0, x, y, r0, (int)min((long)DP_R, n_rows - r0));
  }
  for (; tile < n_tiles; tile += gridDim.x) {
    const long row0 = tile * DP_R;
    const int rows = (int)min((long)DP_R, n_rows - row0);
    // prefetch the next tile into the other stage
    const long next = tile + gridDim.x;
    if (next < n_tiles && tid == 0) {
      const long r0 = next * DP_R;
      dp_issue_tile(s, stage ^ 1, x, y, r0, (int)min((long)DP_R, n_rows - r0));
    }
    dp_wait(&s.bar[stage], phase[stage]);
    phase[stage] ^= 1u;
    // ---- 0. transpose x into feature-major (zero rows beyond the end of the data) ----------------------------------
    for (int e = tid; e < DP_R * (DP_D0 / 4); e += DP_THREADS) {
      const int r = e >> 2, q = e & 3;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows) v = *reinterpret_cast<const float4*>(&s.xs[stage][r * DP_D0 + 4 * q]);
      s.xt[(4 * q + 0) * DP_RS + r] = v.x;
      s.xt[(4 * q + 1) * DP_RS + r] = v.y;
      s.xt[(4 * q + 2) * DP_RS + r] = v.z;
      s.xt[(4 * q + 3) * DP_RS + r] = v.w;
    }
    __syncthreads();
    // ---- 1. H1 = sigmoid(X W0^T + b0) ------------------------------------------------------------------------------
    {
      float acc[8][4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float b = s.b0[tx + 16 * c];
#pragma unroll
        for (int m = 0; m < 8; ++m) acc[m][c] = b;
      }
      gemm_kmajor<8, 4, DP_D0>(s.xt, DP_RS, s.w0t, DP_H, ty * 8, tx * 4, acc);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float h[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) h[m] = dp_sigmoid(acc[m][c]);
        float* dst = &s.A[(tx + 16 * c) * DP_RS + ty * 8];
        *reinterpret_cast<float4*>(dst) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(h[4], h[5], h[6], h[7]);
      }
    }
    __syncthreads();
    // ---- 2. H2 = sigmoid(H1 W1^T + b1) -----------------------------------------------------------------------------
    {
      float acc[8][4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float b = s.b1[tx + 16 * c];
#pragma unroll
        for (int m = 0; m < 8; ++m) acc[m][c] = b;
      }
      gemm_kmajor<8, 4, DP_H>(s.A, DP_RS, s.w1t, DP_H, ty * 8, tx * 4, acc);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float h[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) h[m] = dp_sigmoid(acc[m][c]);
        float* dst = &s.B[(tx + 16 * c) * DP_RS + ty * 8];
        *reinterpret_cast<float4*>(dst) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(h[4], h[5], h[6], h[7]);
      }
    }
    __syncthreads();
    // ---- 3. head: a = H2 w2 + b2, p = sigmoid(a), log-lik term, delta3 = y - p  (stats/loss.py:2 semantics) ---------
    if (tid < DP_R) {
      const int r = tid;
      float a = s.b2;
#pragma unroll 8
      for (int f = 0; f < DP_H; ++f) a = fmaf(s.B[f * DP_RS + r], s.w2[f], a);
      float d = 0.f;
      if (r < rows) {
        const float yv = (r < (rows & ~3)) ? s.ys[stage][r] : y[row0 + r];
        const float p = 1.0f / (1.0f + expf(-a));
        float term;
        if (yv == 1.0f) term = (p == 1.0f) ? NAN : logf(p);
        else if (yv == 0.0f) term = (p == 0.0f) ? NAN : logf(1.0f - p);
        else term = logf(p) * yv + logf(1.0f - p) * (1.0f - yv);
        d = (p == 0.0f || p == 1.0f) ? NAN : (yv - p);
        ll += (double)term;
        gb2 += (double)d;
      }
      s.d3[r] = d;
    }
    __syncthreads();
    // ---- 4. dW2 partial, Delta2 = delta3 w2 H2 (1 - H2) in place, db1 partial ---------------------------------------
    {
      const float w2f = s.w2[ef];
      float sw = 0.f, sb = 0.f;
      float* row = &s.B[ef * DP_RS + es * 32];
#pragma unroll
      for (int r = 0; r < 32; r += 4) {
        float4 h = *reinterpret_cast<float4*>(row + r);
        const float4 d = *reinterpret_cast<const float4*>(&s.d3[es * 32 + r]);
        sw = fmaf(d.x, h.x, sw); sw = fmaf(d.y, h.y, sw); sw = fmaf(d.z, h.z, sw); sw = fmaf(d.w, h.w, sw);
        h.x = d.x * w2f * (1.f - h.x) * h.x; h.y = d.y * w2f * (1.f - h.y) * h.y;
        h.z = d.z * w2f * (1.f - h.z) * h.z; h.w = d.w * w2f * (1.f - h.w) * h.w;
        sb += (h.x + h.y) + (h.z + h.w);
        *reinterpret_cast<float4*>(row + r) = h;
      }
      gw2 += (double)sw;
      gb1 += (double)sb;
    }
    __syncthreads();
    // ---- 5. dW1 += Delta2^T H1 ; D1 = Delta2 W1 ; Delta1 = D1 H1 (1 - H1) in place ----------------------------------
    {
      float t1[4][4];
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n) t1[m][n] = 0.f;
      gemm_kminor<4, 4, 16>(s.B, s.A, ty, tx, t1);
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n) g1[m][n] += (double)t1[m][n];
      float acc[8][4];
#pragma unroll
      for (int m = 0; m < 8; ++m)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[m][c] = 0.f;
      gemm_kmajor<8, 4, DP_H>(s.B, DP_RS, s.w1, DP_H, ty * 8, tx * 4, acc);
      __syncthreads();   // every read of H1 (dW1) is done before it is overwritten
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float* dst = &s.A[(tx + 16 * c) * DP_RS + ty * 8];
        float4 h0 = *reinterpret_cast<float4*>(dst), h1 = *reinterpret_cast<float4*>(dst + 4);
        h0.x = acc[0][c] * (1.f - h0.x) * h0.x; h0.y = acc[1][c] * (1.f - h0.y) * h0.y;
        h0.z = acc[2][c] * (1.f - h0.z) * h0.z; h0.w = acc[3][c] * (1.f - h0.w) * h0.w;
        h1.x = acc[4][c] * (1.f - h1.x) * h1.x; h1.y = acc[5][c] * (1.f - h1.y) * h1.y;
        h1.z = acc[6][c] * (1.f - h1.z) * h1.z; h1.w = acc[7][c] * (1.f - h1.w) * h1.w;
        *reinterpret_cast<float4*>(dst) = h0;
        *reinterpret_cast<float4*>(dst + 4) = h1;
      }
    }
    __syncthreads();
    // ---- 6. dW0 += Delta1^T X ; db0 partial -------------------------------------------------------------------------
    {
      float t0[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
#pragma unroll 2
      for (int r = 0; r < DP_R; r += 4) {
        float4 a[2], b[2];
        a[0] = *reinterpret_cast<const float4*>(&s.A[o5 * DP_RS + r]);
        a[1] = *reinterpret_cast<const float4*>(&s.A[(o5 + 32) * DP_RS + r]);
        b[0] = *reinterpret_cast<const float4*>(&s.xt[j5 * DP_RS + r]);
        b[1] = *reinterpret_cast<const float4*>(&s.xt[(j5 + 8) * DP_RS + r]);
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
          for (int n = 0; n < 2; ++n) {
            t0[m][n] = fmaf(a[m].x, b[n].x, t0[m][n]); t0[m][n] = fmaf(a[m].y, b[n].y, t0[m][n]);
            t0[m][n] = fmaf(a[m].z, b[n].z, t0[m][n]); t0[m][n] = fmaf(a[m].w, b[n].w, t0[m][n]);
          }
      }
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int n = 0; n < 2; ++n) g0[m][n] += (double)t0[m][n];
      float sb = 0.f;
      const float* row = &s.A[ef * DP_RS + es * 32];
#pragma unroll
      for (int r = 0; r < 32; r += 4) {
        const float4 v = *reinterpret_cast<const float4*>(row + r);
        sb += (v.x + v.y) + (v.z + v.w);
      }
      gb0 += (double)sb;
    }
    __syncthreads();
    stage ^= 1;
  }

  // ---- write this CTA's partial sums ---------------------------------------------------------------------------------
  double* out = partials + (size_t)blockIdx.x * (DP_P + 1);
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int n = 0; n < 4; ++n) out[1 + DP_OFF_W1 + (ty + 16 * m) * DP_H + (tx + 16 * n)] = g1[m][n];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < 2; ++n) out[1 + (o5 + 32 * m) * DP_D0 + (j5 + 8 * n)] = g0[m][n];
  // per-feature sums held by 4 threads (ef, es = 0..3)
  auto quad_sum = [&](double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
  };
  const double sb1 = quad_sum(gb1), sb0 = quad_sum(gb0), sw2 = quad_sum(gw2);
  if (es == 0) {
    out[1 + DP_OFF_B1 + ef] = sb1;
    out[1 + DP_OFF_B0 + ef] = sb0;
    out[1 + DP_OFF_W2 + ef] = sw2;
  }
  // log-likelihood and db2: block reduction in a fixed order
  s.red[tid] = ll;
  __syncthreads();
  if (tid == 0) { double t = 0.0; for (int i = 0; i < DP_R; ++i) t += s.red[i]; out[0] = t; }
  __syncthreads();
  s.red[tid] = gb2;
  __syncthreads();
  if (tid == 0) { double t = 0.0; for (int i = 0; i < DP_R; ++i) t += s.red[i]; out[1 + DP_OFF_B2] = t; }
}

// out[e] = sum over CTAs of partials[cta][e], fixed order
__global__ void dp_reduce_kernel(const double* __restrict__ partials, int n_parts, double* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e > DP_P) return;
  double t = 0.0;
  for (int c = 0; c < n_parts; ++c) t += partials[(size_t)c * (DP_P + 1) + e];
  out[e] = t;
}

// ---- replicated-state HMC pieces (P = 5313 vectors; one CTA) ------------------------------------------------------------
// finish an evaluation: sums = all-reduced [loglik, dloglik]; adds the Normal prior (bayesian_model.py:46-50) and the
// temperature; writes target (fp64 scalar) and gradient (fp32 vector).
__global__ void dp_finish_kernel(const double* __restrict__ sums, const float* __restrict__ theta,
                                 const float* __restrict__ ploc, const float* __restrict__ pscale, int has_temp,
                                 double temp, double* __restrict__ target, float* __restrict__ grad) {
  __shared__ double red[DP_THREADS];
  double lp = 0.0;
  for (int j = threadIdx.x; j < DP_P; j += blockDim.x) {
    const double sc = (double)pscale[j], dd = (double)theta[j] - (double)ploc[j];
    lp += -(dd * dd) / (2.0 * sc * sc) - log(sc) - kLogSqrt2PiD;
    double g = sums[1 + j] - dd / (sc * sc);
    if (has_temp) g *= temp;
    grad[j] = (float)g;
  }
  red[threadIdx.x] = lp;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < blockDim.x; ++i) t += red[i];
    double ll = sums[0];
    if (has_temp) { ll *= temp; t *= temp; }
    target[0] = ll + t;
  }
}

// momentum draw + kinetic energy + first half step: p = z + step/2 grad_cur; theta_p = theta_cur + step p   (hmc.py:105,110)
__global__ void dp_hmc_begin_kernel(const float* __restrict__ theta_cur, const float* __restrict__ grad_cur, float step,
                                    RngKey key, uint32_t iter, const float* __restrict__ z_tape, float* __restrict__ p,
                                    float* __restrict__ theta_p, double* __restrict__ kin0) {
  __shared__ double red[DP_THREADS];
  double k = 0.0;
  for (int j0 = threadIdx.x * 4; j0 < DP_P; j0 += blockDim.x * 4) {
    float z[4];
    if (z_tape) {
      for (int c = 0; c < 4; ++c) z[c] = (j0 + c < DP_P) ? z_tape[j0 + c] : 0.f;
    } else {
      U4 w = philox4x32_10(U4{(uint32_t)(j0 / 4), iter, 0u, 0u}, key.k0, key.k1);
      box_muller<float>(Uni<float>::from(w.x), Uni<float>::from(w.y), &z[0], &z[1]);
      box_muller<float>(Uni<float>::from(w.z), Uni<float>::from(w.w), &z[2], &z[3]);
    }
    for (int c = 0; c < 4; ++c) {
      const int j = j0 + c;
      if (j < DP_P) {
        k += (double)z[c] * (double)z[c];
        const float pj = fmaf(0.5f * step, grad_cur[j], z[c]);
        p[j] = pj;
        theta_p[j] = fmaf(step, pj, theta_cur[j]);
      }
    }
  }
  red[threadIdx.x] = k;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0.0; for (int i = 0; i < blockDim.x; ++i) t += red[i]; kin0[0] = 0.5 * t; }
}

// after an evaluation at theta_p: p += w grad_p; if not last: theta_p += step p; if last: kinetic energy   (hmc.py:113-119)
__global__ void dp_hmc_step_kernel(const float* __restrict__ grad_p, float step, int last, float* __restrict__ p,
                                   float* __restrict__ theta_p, double* __restrict__ kin1) {
  __shared__ double red[DP_THREADS];
  const float w = last ? 0.5f * step : step;
  double k = 0.0;
  for (int j = threadIdx.x; j < DP_P; j += blockDim.x) {
    const float pj = fmaf(w, grad_p[j], p[j]);
    p[j] = pj;
    if (!last) theta_p[j] = fmaf(step, pj, theta_p[j]);
    k += (double)pj * (double)pj;
  }
  if (last) {
    red[threadIdx.x] = k;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0.0; for (int i = 0; i < blockDim.x; ++i) t += red[i]; kin1[0] = 0.5 * t; }
  }
}

// accept test in linear space (hmc.py:143-148) and state commit; optional sample write-out
__global__ void dp_hmc_accept_kernel(float* __restrict__ theta_cur, float* __restrict__ grad_cur,
                                     double* __restrict__ target_cur, const float* __restrict__ theta_p,
                                     const float* __restrict__ grad_p, const double* __restrict__ target_p,
                                     const double* __restrict__ kin0, const double* __restrict__ kin1, RngKey key,
                                     uint32_t iter, const float* __restrict__ u_tape, float* __restrict__ out_sample,
                                     double* __restrict__ out_target, uint8_t* __restrict__ out_acc,
                                     uint32_t* __restrict__ acc_count) {
  __shared__ int acc_s;
  if (threadIdx.x == 0) {
    const double h_cur = -target_cur[0] + kin0[0], h_prop = -target_p[0] + kin1[0];
    double rate = exp(h_cur - h_prop);
    rate = (rate > 1.0) ? 1.0 : rate;
    const float u = u_tape ? u_tape[0] : philox_uniform<float>(key, 0u, iter);
    acc_s = ((double)u < rate) ? 1 : 0;
  }
  __syncthreads();
  const int acc = acc_s;
  for (int j = threadIdx.x; j < DP_P; j += blockDim.x) {
    if (acc) { theta_cur[j] = theta_p[j]; grad_cur[j] = grad_p[j]; }
    if (out_sample) out_sample[j] = acc ? theta_p[j] : theta_cur[j];
  }
  if (threadIdx.x == 0) {
    if (acc) target_cur[0] = target_p[0];
    if (out_target) out_target[0] = acc ? target_p[0] : target_cur[0];
    if (out_acc) out_acc[0] = (uint8_t)acc;
    if (acc_count) acc_count[0] += (uint32_t)acc;
  }
}


// ---- fused "post" step of one evaluation (device code in datapar_post.cuh) ------------------------------------------------
// Scalars (target, kinetic energy) are per-CTA partials folded in CTA order by the last CTA to finish (ticket counter).
__global__ void __launch_bounds__(DP_POST_THREADS)
dp_post_kernel(const double* __restrict__ partials, int n_parts, DpExchange xc, const float* __restrict__ theta,
               const float* __restrict__ ploc, const float* __restrict__ pscale, int has_temp, double temp,
               float* __restrict__ grad_out, double* __restrict__ target_out, int step_mode, float step,
               float* __restrict__ mom, float* __restrict__ theta_p, double* __restrict__ kin_out, double* __restrict__ scratch,
               int* __restrict__ status) {
  __shared__ double red[DP_POST_THREADS / 32];
  __shared__ int last_s;
  const int tid = threadIdx.x, cta = blockIdx.x;
  dp_post_entries(partials, n_parts, xc, xc.seq, theta, ploc, pscale, has_temp, temp, grad_out, step_mode, step, mom, theta_p,
                  scratch, status, red, cta, tid);
  if (tid == 0) {
    __threadfence();
    const unsigned long long t = atomicAdd(reinterpret_cast<unsigned long long*>(&scratch[DP_SCRATCH_TICKET]), 1ull);
    last_s = (t == (unsigned long long)(gridDim.x - 1));
  }
  __syncthreads();
  if (last_s && tid == 0) {
    __threadfence();
    double target, kin;
    dp_post_scalars(scratch, has_temp, temp, target, kin);
    target_out[0] = target;
    if (step_mode == 2) kin_out[0] = kin;
    *reinterpret_cast<unsigned long long*>(&scratch[DP_SCRATCH_TICKET]) = 0ull;
  }
}

}  // namespace eb

using namespace eb;

extern "C" {

int eeyore_b200_set_error_(int code, const char* msg);

static int dp_cuda(cudaError_t e, const char* where) {
  if (e == cudaSuccess) return EEYORE_B200_OK;
  std::string m = std::string(where) + ": " + cudaGetErrorString(e);
  return eeyore_b200_set_error_(EEYORE_B200_ECUDA, m.c_str());
}

int eeyore_b200_dp_num_params(void) { return DP_P; }

int64_t eeyore_b200_dp_workspace_bytes(void) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return (int64_t)sizeof(double) * sms * (DP_P + 1);
}

int eeyore_b200_dp_loglik_grad_ffma(const void* theta, const void* x, const void* y, int64_t n_rows, void* out_sums,
                               void* workspace, void* stream) {
  if (!theta || !x || !y || !out_sums || n_rows < 1) return eeyore_b200_set_error_(EEYORE_B200_EINVAL, "dp_loglik_grad_ffma: bad argument");
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15))
    return eeyore_b200_set_error_(EEYORE_B200_EINVAL, "dp_loglik_grad: x and y must be 16-byte aligned (TMA bulk copy)");
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long n_tiles = (n_rows + DP_R - 1) / DP_R;
  const int grid = (int)(n_tiles < sms ? n_tiles : sms);
  double* partials = (double*)workspace;     // caller-owned [SMs, P + 1] doubles, or NULL: stream-ordered temporary
  cudaError_t e;
  if (!workspace) {
    e = cudaMallocAsync((void**)&partials, sizeof(double) * (size_t)grid * (DP_P + 1), st);
    if (e != cudaSuccess) return dp_cuda(e, "dp_loglik_grad(alloc)");
  }
  static bool attr_set = false;
  if (!attr_set) {
    e = cudaFuncSetAttribute(dp_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DpSmem));
    if (e != cudaSuccess) return dp_cuda(e, "dp_loglik_grad(attr)");
    attr_set = true;
  }
  dp_eval_kernel<<<grid, DP_THREADS, sizeof(DpSmem), st>>>((const float*)theta, (const float*)x, (const float*)y,
                                                           (long)n_rows, partials);
  dp_reduce_kernel<<<(DP_P + 1 + 255) / 256, 256, 0, st>>>(partials, grid, (double*)out_sums);
  e = cudaGetLastError();
  if (!workspace) cudaFreeAsync(partials, st);
  return dp_cuda(e, "dp_loglik_grad");
}

int eeyore_b200_dp_finish(const void* sums, const void* theta, const void* prior_loc, const void* prior_scale,
                          int has_temperature, double temperature, void* out_target, void* out_grad, void* stream) {
  if (!sums || !theta || !prior_loc || !prior_scale || !out_target || !out_grad)
    return eeyore_b200_set_error_(EEYORE_B200_EINVAL, "dp_finish: null argument");
  dp_finish_kernel<<<1, DP_THREADS, 0, (cudaStream_t)stream>>>((const double*)sums, (const float*)theta, (const float*)prior_loc,
                                                               (const float*)prior_scale, has_temperature, temperature,
                                                               (double*)out_target, (float*)out_grad);
  return dp_cuda(cudaGetLastError(), "dp_finish");
}

int eeyore_b200_dp_hmc_begin(const void* theta_cur, const void* grad_cur, double step, uint64_t seed, uint64_t iter,
                             const void* z_tape, void* momentum, void* theta_prop, void* kin0, void* stream) {
  RngKey key{(uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32)};
  dp_hmc_begin_kernel<<<1, DP_THREADS, 0, (cudaStream_t)stream>>>((const float*)theta_cur, (const float*)grad_cur, (float)step, key,
                                                                  (uint32_t)iter, (const float*)z_tape, (float*)momentum,
                                                                  (float*)theta_prop, (double*)kin0);
  return dp_cuda(cudaGetLastError(), "dp_hmc_begin");
}

int eeyore_b200_dp_hmc_step(const void* grad_prop, double step, int last, void* momentum, void* theta_prop, void* kin1,
                            void* stream) {
  dp_hmc_step_kernel<<<1, DP_THREADS, 0, (cudaStream_t)stream>>>((const float*)grad_prop, (float)step, last, (float*)momentum,
                                                                 (float*)theta_prop, (double*)kin1);
  return dp_cuda(cudaGetLastError(), "dp_hmc_step");
}

int eeyore_b200_dp_hmc_accept(void* theta_cur, void* grad_cur, void* target_cur, const void* theta_prop,
                              const void* grad_prop, const void* target_prop, const void* kin0, const void* kin1,
                              uint64_t seed, uint64_t iter, const void* u_tape, void* out_sample, void* out_target,
                              uint8_t* out_accepted, uint32_t* accept_count, void* stream) {
  RngKey key{(uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32)};
  dp_hmc_accept_kernel<<<1, DP_THREADS, 0, (cudaStream_t)stream>>>(
      (float*)theta_cur, (float*)grad_cur, (double*)target_cur, (const float*)theta_prop, (const float*)grad_prop,
      (const double*)target_prop, (const double*)kin0, (const double*)kin1, key, (uint32_t)iter, (const float*)u_tape,
      (float*)out_sample, (double*)out_target, out_accepted, accept_count);
  return dp_cuda(cudaGetLastError(), "dp_hmc_accept");
}

/* ---- peer exchange area of the data-sharded path (CUDA IPC; one per rank) ---------------------------------------------- */
int64_t eeyore_b200_dp_exchange_bytes(void) {
  return (int64_t)(2 * DP_XMAXW * DP_XSLOT * sizeof(double) + 2 * DP_XMAXW * 32 * sizeof(unsigned long long) + DP_SCRATCH_LEN * sizeof(double));
}

int64_t eeyore_b200_dp_scratch_len(void) { return DP_SCRATCH_LEN; }

int64_t eeyore_b200_dp_exchange_scratch_offset(void) {
  return (int64_t)(2 * DP_XMAXW * DP_XSLOT * sizeof(double) + 2 * DP_XMAXW * 32 * sizeof(unsigned long long));
}

int eeyore_b200_dp_exchange_create(void** out_base, void* out_handle64) {
  if (!out_base || !out_handle64) return eeyore_b200_set_error_(EEYORE_B200_EINVAL, "dp_exchange_create: null argument");
  void* base = nullptr;
  cudaError_t e = cudaMalloc(&base, (size_t)eeyore_b200_dp_exchange_bytes());
  if (e != cudaSuccess) return dp_cuda(e, "dp_exchange_create(malloc)");
  e = cudaMemset(base, 0, (size_t)eeyore_b200_dp_exchange_bytes());
  if (e != cudaSuccess) return dp_cuda(e, "dp_exchange_create(memset)");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(out_handle64), base);
  if (e != cudaSuccess) return dp_cuda(e, "dp_exchange_create(ipc handle)");
  e = cudaDeviceSynchronize();
  *out_base = base;
  return dp_cuda(e, "dp_exchange_create");
}

int eeyore_b200_dp_exchange_open(const void* handle64, void** out_ptr) {
  if (!handle64 || !out_ptr) return eeyore_b200_set_error_(EEYORE_B200_EINVAL, "dp_exchange_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  return dp_cuda(cudaIpcOpenMemHandle(out_ptr, h, cudaIpcMemLazyEnablePeerAccess), "dp_exchange_open");
}

int eeyore_b200_dp_exchange_close(void* peer_ptr) { return dp_cuda(cudaIpcCloseMemHandle(peer_ptr), "dp_exchange_close"); }
int eeyore_b200_dp_exchange_destroy(void* base) { return dp_cuda(cudaFree(base), "dp_exchange_destroy"); }

int eeyore_b200_dp_post(const void* workspace, int n_parts, int world, int rank, uint64_t seq, void* const* peer_bases,
                        void* local_scratch, const void* theta, const void* prior_loc, const void* prior_scale,
                        int has_temperature, double temperature, void* out_grad, void* out_target, int step_mode,
                        double step, void* momentum, void* theta_prop, void* kin1, int32_t* status, void* stream) {
  if (!workspace || !theta || !prior_loc || !prior_scale || !out_grad || !out_target || !local_scratch || !status)
    return eeyore_b200_set_error_(EEYORE_B200_EINVAL, "dp_post: null argument");
  if (world < 1 || world > DP_XMAXW || rank < 0 || rank >= world || (world > 1 && (!peer_bases || seq == 0)))
    return eeyore_b200_set_error_(EEYORE_B200_EINVAL, "dp_post: bad world / rank / sequence number");
  if (step_mode && (!momentum || !theta_prop || !kin1)) return eeyore_b200_set_error_(EEYORE_B200_EINVAL, "dp_post: leapfrog buffers missing");
  DpExchange xc{};
  xc.world = world;
  xc.rank = rank;
  xc.seq = seq;
  for (int p = 0; p < world && world > 1; ++p) {
    xc.inbox[p] = reinterpret_cast<double*>(peer_bases[p]);
    xc.flags[p] = reinterpret_cast<unsigned long long*>(xc.inbox[p] + 2 * DP_XMAXW * DP_XSLOT);
  }
  dp_post_kernel<<<DP_POST_CTAS, DP_POST_THREADS, 0, (cudaStream_t)stream>>>(
      (const double*)workspace, n_parts, xc, (const float*)theta, (const float*)prior_loc, (const float*)prior_scale,
      has_temperature, temperature, (float*)out_grad, (double*)out_target, step_mode, (float)step, (float*)momentum,
      (float*)theta_prop, (double*)kin1, (double*)local_scratch, status);
  return dp_cuda(cudaGetLastError(), "dp_post");
}

/* number of per-CTA partial-sum rows the evaluation kernel writes for n_rows rows (the n_parts of dp_post) */
int eeyore_b200_dp_num_parts(int64_t n_rows) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long n_tiles = (n_rows + DP_R - 1) / DP_R;
  return (int)(n_tiles < sms ? n_tiles : sms);
}

}  // extern "C"
