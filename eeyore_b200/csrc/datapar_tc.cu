// Tensor-core (tcgen05 + TMEM) version of the data-parallel log-likelihood + gradient kernel of BASELINE config 5
// (MLP 16-64-64-1, fp32, millions of rows, one parameter vector).  Same contract and same output as
// dp_eval_kernel in datapar.cu, which it replaces on the product path:
//   eeyore/models/mlp.py:45-50 (forward), eeyore/stats/loss.py:1-11 (naive BCE on probabilities),
//   eeyore/models/bayesian_model.py:30-35 (log_lik), eeyore/models/log_target_model.py:15-23 (gradient).
//
// Why tensor cores here and nowhere else: the two hidden layers are dense 64-wide contractions over 128-row tiles
// (29,056 FLOP per row); the chain-batched kernels have no such contraction.
//
// fp32 parity on 16-bit tensor cores: operands are split exactly into 16-bit pieces and a product A B is the sum of the
// leading piece products, accumulated in fp32 in TMEM.  The B pieces are laid out side by side along N, so one A piece is
// read once for all the B pieces it meets (one MMA per A piece, one 64-column accumulator group per B piece).
//   Every operand is two fp16 pieces (11 + 11 bits, exact residual) after an exact power-of-two scaling that brings its
//   largest possible magnitude just below 2^14: W0, W1 from max |W| (prologue), x from the shard's max |x| (a device scalar
//   supplied by the caller or computed by a reduction kernel), H1 by 2^10, Delta2 from |Delta2| <= max|w2| / 4, Delta1 from
//   |Delta1| <= 4 max|w2| max|W1|.  A product is A1 B1 + A1 B2 + A2 B1 (dropped term <= 2^-22) = two MMAs (A1 x [B1 B2],
//   A2 x [B1]); the scales are undone in the epilogues (folded into an existing multiply).  The first version used three
//   bf16 pieces and six products (no scaling needed): same accuracy, 1.75x the tensor time and 1.5x the operand bytes.
//   Emulated on the CPU and measured on the GPU: gradient within 3e-7 of the fp64 oracle (bar 1e-5).
//
// Two tiles are in flight per CTA (contexts A and B with their own operand buffers, accumulator region and mbarriers): the
// phases alternate A, B, A, B ..., so the MMAs of one tile run on the tensor pipe under the epilogue of the other.
// Per 128-row tile (persistent CTAs, one per SM, 256 threads; thread = (row, half of the 64 features)):
//   P0  x tile (coalesced loads, prefetched one pair of tiles ahead) -> pieces      MMA1  Z1 = X W0^T
//   P1  H1 = sigmoid(Z1 + b0) -> pieces (smem)                                       MMA2  Z2 = H1 W1^T
//   P2  H2, head, log-lik, delta3, dW2 (warp butterfly), Delta2 -> pieces           MMA3  D1 = Delta2 W1
//                                                                                    MMA4  [db1 dW1] = Delta2^T [1 H1]
//   P3  Delta1 = D1 H1 (1 - H1) -> pieces (over Delta2, after MMA4; H1 re-read from its pieces)  MMA5  [db0 dW0] = Delta1^T [1 X]
//   P4  every TC_FLUSH pairs of tiles: weight-gradient sums TMEM -> (sum, error) fp32 register pairs (fp32 accumulation in
//       TMEM spans <= 512 rows)
// (A ninth, issue-only warp was measured slower: three warps on one scheduler cap the kernel at 168 registers -> spills.)
// One shared-memory copy of each activation serves both orientations: the SWIZZLE_NONE core-matrix layout of tc05.cuh
// is a K-major operand for the forward / back-propagation GEMMs and an MN-major operand for the weight-gradient GEMMs.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string>
#include "philox.cuh"
#include "datapar.cuh"
#include "datapar_post.cuh"
#include "tc05.cuh"
#include "../../include/eeyore_b200.h"

namespace eb {

using namespace tc;

constexpr uint32_t TC_CS = 2048;                 // chunk stride of [128 x C] activation buffers (128 rows * 16 B)
constexpr uint32_t TC_ACT = 8 * TC_CS;           // one bf16 piece of a [128 x 64] activation: 16 KB
constexpr uint32_t TC_XP = 2 * TC_CS;            // one piece of the [128 x 16] x tile: 4 KB
constexpr uint32_t TC_WCS2 = 128 * 16;           // chunk stride of the 2-piece-stacked weight buffers ([128 x K])
#ifndef EB_TC_FMA_EVERY
#define EB_TC_FMA_EVERY 2
#endif
// every EB_TC_FMA_EVERY-th sigmoid takes its reciprocal on the FMA pipe instead of the MUFU unit (0 = none)
#define TC_FMA_RCP(j) (EB_TC_FMA_EVERY > 0 && ((j) % (EB_TC_FMA_EVERY > 0 ? EB_TC_FMA_EVERY : 1)) == 1)
constexpr float TC_NL2E = -1.4426950408889634f;  // -log2 e
constexpr float TC_SH = 1024.f;                  // scale of H1 before the fp16 split (keeps the low piece normal)
// TMEM columns (fp32)
constexpr uint32_t TM_Z = 0;                     // per context (c * 128): columns 0..63 Z1, then Z2, then D1; 64..127 sh^2 H1 (1 - H1)
constexpr uint32_t TM_W1 = 256;                  // [ones(8) | dW1 (2 x 64)] on lanes 16q..16q+15, accumulated over TC_FLUSH tiles
constexpr uint32_t TM_W0 = 392;                  // [ones(8) | dW0 (2 x 16)]     (432 of 512 columns)
constexpr int TC_FLUSH = 2;                      // PAIRS of tiles between two folds (4 tiles)
#ifndef EB_TC_CG
#define EB_TC_CG 2
#endif
constexpr int TC_CG = EB_TC_CG;                 // column groups: thread = (row, group of TC_FW of the 64 hidden units)
// 4 was round 1's 16-warp experiment (same speed at 128 registers); since round 2 the persistent run kernel and the register
// split with the issue warp group assume the 8 epilogue warps of TC_CG = 2.
static_assert(TC_CG == 2, "EB_TC_CG: only 2 column groups (256 epilogue threads) are supported");
constexpr int TC_FW = DP_H / TC_CG;              // hidden units per thread
constexpr int TC_XW = DP_D0 / TC_CG;             // dW0 columns per thread
constexpr int TC_THREADS = 128 * TC_CG;          // epilogue threads (warps 0 .. 4 TC_CG - 1)
constexpr int TC_LAUNCH = TC_THREADS + 128;      // + one warp group whose first warp only issues the MMAs (see tc_issuer_body)
constexpr int TC_REG_EPI = 232, TC_REG_ISSUE = 40;   // setmaxnreg: 2 x 232 + 40 = 3 x 168 (the launch bound of 384 threads)

struct TcCtx {                                   // operand buffers of one tile in flight
  alignas(1024) uint16_t ones_h[1024];           // 2 KB of fp16 1.0: the N-chunk in front of the H1 pieces
  alignas(16) uint16_t h1[2 * TC_ACT / 2];       // H1 pieces (fp16 x 2, scaled by TC_SH)
  alignas(16) uint16_t ones_x[1024];             // 2 KB of fp16 1.0: the N-chunk in front of the x pieces
  alignas(16) uint16_t xp[2 * TC_XP / 2];        // x pieces (fp16 x 2, scaled)
  alignas(16) uint16_t dl[2 * TC_ACT / 2];       // Delta2 pieces, then Delta1 pieces (fp16 x 2, scaled)
};

struct TcSmem {
  TcCtx ctx[2];
  alignas(16) uint16_t w0s[2 * TC_WCS2 / 2];     // fp16 x 2, scaled: rows 64 p + o, cols j      (B of MMA1)
  alignas(16) uint16_t w1a[8 * TC_WCS2 / 2];     // fp16 x 2, scaled: rows 64 p + o, cols i      (B of MMA2)
  alignas(16) uint16_t w1b[8 * TC_WCS2 / 2];     // fp16 x 2, scaled: rows 64 p + i, cols o      (B of MMA3)
  alignas(16) float b0[DP_H], b1[DP_H], w2[DP_H], w2d[DP_H];   // w2d = w2 * (scale of Delta2)
  alignas(16) float exch[2][TC_CG][DP_R];          // per context: the head's partial sums of the column groups
  alignas(8) unsigned long long bar[2][6];       // per context, 1..5: MMA groups (completion, tcgen05.commit)
  alignas(8) unsigned long long ready[2];        // per context: the epilogue warps have finished the current phase (8 arrivals)
  float b2;
  float red_max[3][TC_THREADS / 32];             // prologue: max |W1|, max |w2|, max |W0| per warp
  uint32_t tmem_base;
};
static_assert(sizeof(TcSmem) <= 227 * 1024, "shared memory budget");

#ifdef DP_TC_PROFILE
__device__ unsigned long long dp_tc_prof[24];
#define TC_STAMP(i)                                                        \
  do {                                                                     \
    if (blockIdx.x == 0 && tid == 32) {                                    \
      const long long now_ = clock64();                                    \
      dp_tc_prof[i] += (unsigned long long)(now_ - last_);                 \
      last_ = now_;                                                        \
    }                                                                      \
  } while (0)
#else
#define TC_STAMP(i)
#endif

// SC / (1 + 2^zl), zl = -z log2 e (the factor rides on the bias and on the scale of the pre-activation), SC = 1 or TC_SH: two MUFU ops (ex2, rcp; both within 2 ulp) and no slow path -- exp overflow gives
// exactly 0.  The power-of-two scale of the fp16 split rides on the reciprocal's argument: rcp((1 + e) / SC) = SC rcp(1 + e)
// bit for bit (one FFMA instead of FADD + FMUL).
template <bool SCALED>
__device__ __forceinline__ float tc_sigmoid(float zl) {
  constexpr float c = SCALED ? 1.0f / TC_SH : 1.0f;
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(zl));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(e, c, c)));
  return r;
}

// The same value with the reciprocal on the FMA pipe (integer seed, error <= 12 %, three Newton steps -> 4e-8).  With two
// tiles in flight the epilogues are the critical path and the MUFU pipe (two ops per sigmoid) is their busiest unit: every
// other unit takes this route (measured: 0 % 0.539 ms, 50 % 0.525 ms, 75 % 0.534 ms per 2M rows).
template <bool SCALED>
__device__ __forceinline__ float tc_sigmoid_fma(float zl) {
  constexpr float c = SCALED ? 1.0f / TC_SH : 1.0f;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(zl));
  const float d = fminf(fmaf(e, c, c), 1.0e30f);
  float r = __uint_as_float(0x7EF311C7u - __float_as_uint(d));
  r = r * fmaf(-d, r, 2.0f);
  r = r * fmaf(-d, r, 2.0f);
  r = r * fmaf(-d, r, 2.0f);
  return r;
}
// fp16 pair {hi half: b, lo half: a} (round to nearest even) and back
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
  uint32_t d;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t h) {
  float2 r;
  asm("{\n\t.reg .f16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tcvt.f32.f16 %0, lo;\n\tcvt.f32.f16 %1, hi;\n\t}" : "=f"(r.x), "=f"(r.y) : "r"(h));
  return r;
}
// x0 - lo(h), x1 - hi(h) for h = pack_f16x2(x0, x1): Blackwell's mixed-precision FMA (fp16 x fp16 + fp32 -> fp32, SASS FHFMA)
// takes the packed halves as they are -- no conversion back to fp32 -- and the difference is exact.
__device__ __forceinline__ void resid_f16x2(uint32_t h, float x0, float x1, float& r0, float& r1) {
  asm("{\n\t.reg .f16 lo, hi, m1;\n\tmov.b32 {lo, hi}, %2;\n\tmov.b16 m1, 0xBC00;\n\t"
      "fma.rn.f32.f16 %0, lo, m1, %3;\n\tfma.rn.f32.f16 %1, hi, m1, %4;\n\t}"
      : "=f"(r0), "=f"(r1) : "r"(h), "f"(x0), "f"(x1));
}
// float(a) + float(b) per half of two fp16 pairs (one conversion + one mixed-precision add, SASS FHADD, per value)
__device__ __forceinline__ float2 sum_f16x2(uint32_t a, uint32_t b) {
  float2 r;
  asm("{\n\t.reg .f16 a0, a1, b0, b1;\n\t.reg .f32 t0, t1;\n\tmov.b32 {a0, a1}, %2;\n\tmov.b32 {b0, b1}, %3;\n\t"
      "cvt.f32.f16 t0, a0;\n\tcvt.f32.f16 t1, a1;\n\tadd.rn.f32.f16 %0, b0, t0;\n\tadd.rn.f32.f16 %1, b1, t1;\n\t}"
      : "=f"(r.x), "=f"(r.y) : "r"(a), "r"(b));
  return r;
}
// 8 consecutive fp32 values (already scaled) -> two 16-byte rows of fp16 pieces: v = p1 + p2 + O(2^-22 v) (exact residual)
__device__ __forceinline__ void split2h(const float* v, uint4& p1, uint4& p2) {
  uint32_t a[4], b[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float x0 = v[2 * j], x1 = v[2 * j + 1];
    const uint32_t h = pack_f16x2(x0, x1);
    float r0, r1;
    resid_f16x2(h, x0, x1, r0, r1);
    a[j] = h;
    b[j] = pack_f16x2(r0, r1);
  }
  p1 = make_uint4(a[0], a[1], a[2], a[3]);
  p2 = make_uint4(b[0], b[1], b[2], b[3]);
}
__device__ __forceinline__ void split2h_scalar(float x, uint16_t& p1, uint16_t& p2) {
  const uint32_t h = pack_f16x2(x, 0.f);
  const float r = x - unpack_f16x2(h).x;
  p1 = (uint16_t)h;
  p2 = (uint16_t)pack_f16x2(r, 0.f);
}
// TC_FW features of one row (already scaled), starting at chunk `chunk0` -> the two fp16 piece buffers
__device__ __forceinline__ void store_pieces32h(unsigned char* base, int chunk0, int r, const float* v) {
  unsigned char* p = base + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u + (uint32_t)chunk0 * TC_CS;
#pragma unroll
  for (int c = 0; c < TC_FW / 8; ++c) {
    uint4 p1, p2;
    split2h(v + 8 * c, p1, p2);
    *reinterpret_cast<uint4*>(p + c * TC_CS) = p1;
    *reinterpret_cast<uint4*>(p + c * TC_CS + TC_ACT) = p2;
  }
}
// 2^e (|e| < 127) and the exponent of the power of two that brings `vmax` into [2^13, 2^14)
__device__ __forceinline__ float pow2i(int e) { return __int_as_float((e + 127) << 23); }
__device__ __forceinline__ int scale_exp(float vmax) {
  if (!(vmax > 0.f) || !(vmax < 3.0e38f)) return 0;       // all zero, or inf / nan (which propagate anyway)
  const int e = 13 - (((__float_as_int(vmax) >> 23) & 0xff) - 127);
  return e < -40 ? -40 : (e > 40 ? 40 : e);
}

// v[j] = accumulator columns taddr + j, j < TC_FW
__device__ __forceinline__ void load_acc(uint32_t taddr, float* v) {
  if constexpr (TC_FW == 32) {
    uint32_t a[32];
    tmem_ld32(taddr, a);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(a[j]);
  } else {
    uint32_t a[16];
    tmem_ld16(taddr, a);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(a[j]);
  }
}
// v[j] = sum of the two accumulator groups (64 columns apart) of a two-piece product at columns col0 + j, j < 32
__device__ __forceinline__ void load_sum2(uint32_t taddr, float* v) {
  if constexpr (TC_FW == 32) {
    uint32_t a[32], b[32];
    tmem_ld32(taddr, a);
    tmem_ld32(taddr + 64, b);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(b[j]) + __uint_as_float(a[j]);
  } else {
    uint32_t a[16], b[16];
    tmem_ld16(taddr, a);
    tmem_ld16(taddr + 64, b);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(b[j]) + __uint_as_float(a[j]);
  }
}

// hi + lo += x, error-free (Knuth two-sum): a pair of fp32 registers carries ~46 significant bits.  The tile loop uses
// these instead of FP64 adds: ncu showed every isolated DADD of the loop stalling for hundreds of cycles on the FP64 pipe
// (stall_math_pipe_throttle), 8 % of the kernel for a handful of instructions per tile.
__device__ __forceinline__ void acc2(float& hi, float& lo, float x) {
  const float s = hi + x;
  const float bb = s - hi;
  const float e = (hi - (s - bb)) + (x - bb);
  hi = s;
  lo += e;
}

// reduce-scatter over the 32 lanes: on return t[0] of lane l = sum over lanes of their t[l]
template <int W>
__device__ __forceinline__ void butterfly_step(float* t, int lane) {
  const bool up = (lane & W) != 0;
#pragma unroll
  for (int i = 0; i < W; ++i) {
    const float send = up ? t[i] : t[i + W];
    const float keep = up ? t[i + W] : t[i];
    t[i] = keep + __shfl_xor_sync(0xffffffffu, send, W);
  }
}

// one A piece per MMA against the first (NP - a) B pieces; D groups are N_PIECE columns wide
//   a_desc[a]: descriptors of the A pieces;  b_desc: descriptor of [B1 B2 ..] (pieces adjacent along N);  FMT 0 = fp16, 1 = bf16
template <int NP, int FMT, int M, int N_LEAD, int N_PIECE, int A_MN, int B_MN>
__device__ __forceinline__ void mma_product(uint32_t d_tmem, const uint64_t* a_desc, uint64_t b_desc, uint32_t a_step,
                                            uint32_t b_step, int k_steps, uint32_t accumulate = 0u) {
  constexpr uint32_t id0 = idesc_f16kind(M, N_LEAD + NP * N_PIECE, A_MN, B_MN, FMT, FMT);
  constexpr uint32_t id1 = idesc_f16kind(M, N_LEAD + (NP - 1) * N_PIECE, A_MN, B_MN, FMT, FMT);
  constexpr uint32_t id2 = idesc_f16kind(M, N_LEAD + 1 * N_PIECE, A_MN, B_MN, FMT, FMT);
  for (int k = 0; k < k_steps; ++k) {
    const uint64_t bd = desc_advance(b_desc, k * b_step);
    mma_bf16(d_tmem, desc_advance(a_desc[0], k * a_step), bd, id0, (k > 0) ? 1u : accumulate);
    mma_bf16(d_tmem, desc_advance(a_desc[1], k * a_step), bd, id1, 1u);
    if constexpr (NP == 3) mma_bf16(d_tmem, desc_advance(a_desc[2], k * a_step), bd, id2, 1u);
  }
}

__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 5, %0;" ::"n"(TC_THREADS) : "memory"); }   // the epilogue warps
template <int R> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }
__device__ __forceinline__ void mbar_arrive(void* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// End of an epilogue phase of one context: this warp's operand pieces are in shared memory and visible to the async proxy,
// its TMEM reads have completed.  One arrival per warp; nobody waits here -- the issue warp does (tc_issuer_body).
__device__ __forceinline__ void phase_done(unsigned long long* ready, int lane) {
  fence_async_smem();
  fence_before_sync();
  __syncwarp();
  if (lane == 0) mbar_arrive(ready);
}

// A1 B1 + A1 B2 + A2 B1 accumulated into ONE group of N columns (three MMAs per k-step; B2 = B1 + 64 rows = 1024 bytes in the
// core-matrix layout).  The tensor pipe has the slack (a quarter busy) and the epilogue is what bounds the kernel: summing
// the piece products in TMEM instead of two accumulator groups in the epilogue saves a TMEM load and TC_FW additions per
// thread and phase.
template <int M, int N>
__device__ __forceinline__ void mma_product_sum(uint32_t d_tmem, const uint64_t* a_desc, uint64_t b_desc, uint32_t a_step,
                                                uint32_t b_step, int k_steps) {
  constexpr uint32_t id = idesc_f16kind(M, N, 0, 0, 0, 0);
  for (int k = 0; k < k_steps; ++k) {
    const uint64_t b1 = desc_advance(b_desc, k * b_step), b2 = desc_advance(b1, 1024);
    const uint64_t a1 = desc_advance(a_desc[0], k * a_step), a2 = desc_advance(a_desc[1], k * a_step);
    mma_bf16(d_tmem, a1, b1, id, (k > 0) ? 1u : 0u);
    mma_bf16(d_tmem, a1, b2, id, 1u);
    mma_bf16(d_tmem, a2, b1, id, 1u);
  }
}

// One-time set-up of a CTA: the blocks of fp16 ones, the mbarriers, the TMEM allocation.  Returns the TMEM base address.
__device__ __forceinline__ uint32_t tc_setup(TcSmem& s) {
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 1024; e += TC_THREADS) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      s.ctx[c].ones_h[e] = 0x3C00;     // fp16 1.0
      s.ctx[c].ones_x[e] = 0x3C00;
    }
  }
  if (tid == 0) {
#pragma unroll
    for (int b = 0; b < 12; ++b) mbar_init(&s.bar[0][0] + b, 1);
    mbar_init(&s.ready[0], TC_THREADS / 32);
    mbar_init(&s.ready[1], TC_THREADS / 32);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(&s.tmem_base);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  return s.tmem_base;
}

// One evaluation over the rows of this CTA's tiles.  out: this CTA's row of the partial sums ([DP_P + 1] doubles: [0] =
// log-likelihood, [1 + j] = d loglik / d theta_j).  theta is read through L2 (ld.global.cg): inside the persistent HMC kernel
// it was written by other CTAs of the same launch.  `again`: not the first evaluation of this launch -- every MMA of the
// previous one has completed, and the mbarriers (whose phases depend on the tile count) are re-initialised.
__device__ __forceinline__ void tc_eval_body(TcSmem& s, const uint32_t tm, const float* theta, const float* __restrict__ x,
                                             const float* __restrict__ y, const long n_rows, const float* __restrict__ x_absmax,
                                             double* out, const bool again) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, hf = warp >> 2;   // TMEM lane quadrant; which half of the 64 features
  const int r = 32 * q + lane;              // this thread's row of the tile (= TMEM lane)
  const long n_tiles = (n_rows + DP_R - 1) / DP_R;
#ifdef DP_TC_PROFILE
  long long last_ = clock64();
#endif

  // ---- per-evaluation staging: the parameter vector is read ONCE, 21 coalesced L2 loads per thread all in flight together
  // (element e = tid + 256 k; the loop versions of this prologue -- one pass for the maxima, one per matrix for the pieces,
  // each a chain of dependent load latencies -- took 21,000 cycles per evaluation); then the power-of-two scales of the fp16
  // operands from the parameter magnitudes (identical in every CTA and on every rank) and the weights as fp16 pieces.
  constexpr int TV = (DP_P + TC_THREADS - 1) / TC_THREADS;   // 21
  float tv[TV];
#pragma unroll
  for (int k = 0; k < TV; ++k) {
    const int e = tid + TC_THREADS * k;
    tv[k] = e < DP_P ? __ldcg(theta + e) : 0.f;
  }
  {
    float m1 = 0.f, m2 = 0.f, m0 = 0.f;
#pragma unroll
    for (int k = 0; k < TV; ++k) {
      const int e = tid + TC_THREADS * k;
      const float av = fabsf(tv[k]);
      if (e < DP_OFF_B0) m0 = fmaxf(m0, av);
      else if (e >= DP_OFF_W1 && e < DP_OFF_B1) m1 = fmaxf(m1, av);
      else if (e >= DP_OFF_W2 && e < DP_OFF_B2) m2 = fmaxf(m2, av);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
      m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, o));
    }
    if (lane == 0) { s.red_max[0][warp] = m1; s.red_max[1][warp] = m2; s.red_max[2][warp] = m0; }
  }
  epi_sync();
  float w1max = 0.f, w2max = 0.f, w0max = 0.f;
#pragma unroll
  for (int w = 0; w < TC_THREADS / 32; ++w) {
    w1max = fmaxf(w1max, s.red_max[0][w]);
    w2max = fmaxf(w2max, s.red_max[1][w]);
    w0max = fmaxf(w0max, s.red_max[2][w]);
  }
  const int e_w = scale_exp(w1max);                 // W1 * 2^e_w      in [2^13, 2^14)
  const int e_d = scale_exp(0.25f * w2max);         // |Delta2| <= max|w2| / 4:  Delta2 * 2^e_d below 2^14
  const int e_w0 = scale_exp(w0max);
  const int e_x = scale_exp(x_absmax[0]);           // the shard's max |x|
  const int e_d1 = scale_exp(4.0f * w2max * w1max); // |Delta1| <= 64 max|Delta2| max|W1| / 4
  const float s_w = pow2i(e_w), s_d = pow2i(e_d), s_w0 = pow2i(e_w0), s_x = pow2i(e_x);
  const float inv_z1 = pow2i(-e_x) * pow2i(-e_w0) * TC_NL2E;    // -log2 e Z1, Z1 = (X sx)(W0 sw0)^T
  const float inv_w0 = pow2i(-e_d1) * pow2i(-e_x);              // dW0 = (Delta1 sd1)^T (X sx)
  const float inv_b0 = pow2i(-e_d1);                            // db0 = (Delta1 sd1)^T 1
  const float inv_z2 = pow2i(-e_w) * (1.0f / TC_SH) * TC_NL2E;  // -log2 e Z2, Z2 = (H1 sh)(W1 sw)^T
  const float c_d1 = pow2i(-e_w - e_d + e_d1) * (1.0f / (TC_SH * TC_SH));   // Delta1 sd1 = D1' c_d1 (sh - H1')(H1'), primes = scaled
  const float inv_w1 = pow2i(-e_d) * (1.0f / TC_SH);            // dW1 = (Delta2 sd)^T (H1 sh)
  const float inv_b1 = pow2i(-e_d);                             // db1 = (Delta2 sd)^T 1
#pragma unroll
  for (int k = 0; k < TV; ++k) {
    const int e = tid + TC_THREADS * k;
    uint16_t p[2];
    if (e < DP_OFF_B0) {                                  // W0[o][j]                                   (B of MMA1)
      const int o = e / DP_D0, j = e % DP_D0;
      split2h_scalar(tv[k] * s_w0, p[0], p[1]);
#pragma unroll
      for (int c = 0; c < 2; ++c)
        *reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(s.w0s) + cm_off(64 * c + o, j, TC_WCS2)) = p[c];
    } else if (e < DP_OFF_W1) {
      s.b0[e - DP_OFF_B0] = tv[k] * TC_NL2E;                 // biases pre-multiplied by -log2 e (see tc_sigmoid)
    } else if (e < DP_OFF_B1) {                           // W1[o][i]                                   (B of MMA2 and of MMA3)
      const int o = (e - DP_OFF_W1) / DP_H, i = (e - DP_OFF_W1) % DP_H;
      split2h_scalar(tv[k] * s_w, p[0], p[1]);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        *reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(s.w1a) + cm_off(64 * c + o, i, TC_WCS2)) = p[c];
        *reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(s.w1b) + cm_off(64 * c + i, o, TC_WCS2)) = p[c];
      }
    } else if (e < DP_OFF_W2) {
      s.b1[e - DP_OFF_B1] = tv[k] * TC_NL2E;
    } else if (e < DP_OFF_B2) {
      s.w2[e - DP_OFF_W2] = tv[k];
      s.w2d[e - DP_OFF_W2] = tv[k] * s_d;
    } else if (e == DP_OFF_B2) {
      s.b2 = tv[k];
    }
  }
  if (again) {   // the previous evaluation staged its row of sums over the operand buffers, the blocks of ones included
    for (int e = tid; e < 1024; e += TC_THREADS) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        s.ctx[c].ones_h[e] = 0x3C00;
        s.ctx[c].ones_x[e] = 0x3C00;
      }
    }
    if (tid == 0) {
#pragma unroll
      for (int b = 0; b < 12; ++b) { mbar_inval(&s.bar[0][0] + b); mbar_init(&s.bar[0][0] + b, 1); }
#pragma unroll
      for (int c = 0; c < 2; ++c) { mbar_inval(&s.ready[c]); mbar_init(&s.ready[c], TC_THREADS / 32); }
      mbar_fence_init();
    }
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();           // the whole CTA: the issue warp starts from here (tc_issuer_body)
  fence_after_sync();
  const uint32_t tm_lane = tm + ((uint32_t)(32 * q) << 16);   // this warp's lane quadrant

  // weight-gradient sums: lanes 0..15 of every warp own unit o = 16 q + lane (M = 64 accumulator layout)
  float g1[TC_FW], g1e[TC_FW], g0[TC_XW], g0e[TC_XW];   // (sum, error term) pairs, see acc2
#pragma unroll
  for (int i = 0; i < TC_FW; ++i) g1[i] = g1e[i] = 0.f;
#pragma unroll
  for (int i = 0; i < TC_XW; ++i) g0[i] = g0e[i] = 0.f;
  float gb1 = 0.f, gb1e = 0.f, gb0 = 0.f, gb0e = 0.f, gw2 = 0.f, gw2e = 0.f, gb2 = 0.f, gb2e = 0.f, ll = 0.f, lle = 0.f;

  // this thread's 8 features of its row and the row's label for both contexts, loaded one pair of tiles ahead
  float4 xa0 = make_float4(0.f, 0.f, 0.f, 0.f), xb0 = xa0, xa1 = xa0, xb1 = xa0;
  float yn0 = 0.f, yn1 = 0.f;
  auto prefetch_x = [&](long tile, float4& xa, float4& xb, float& yn) {
    const long gr = tile * DP_R + r;
    xa = xb = make_float4(0.f, 0.f, 0.f, 0.f);
    yn = 0.f;
    if (tile < n_tiles && gr < n_rows) {
      if (hf < 2) {                                    // the 16 input features are two 8-wide chunks
        const float4* src = reinterpret_cast<const float4*>(x + gr * DP_D0 + 8 * hf);
        xa = __ldg(src);
        xb = __ldg(src + 1);
      }
      yn = __ldg(y + gr);
    }
  };

  // Fold of the weight-gradient sums (TMEM -> (sum, error) register pairs).  It runs one phase late -- after P1 of the NEXT pair of
  // tiles, when the group's last MMA5 has long completed (waiting for it right behind P3 stalled every warp: 3 % of the samples)
  // and before that pair's P2 hand-over, after which the issue warp may restart the sums (keep = 0).
  bool fold_due = false;
  int fctx = 0;
  uint32_t fpar = 0;
  auto do_fold = [&]() {
      mbar_wait(&s.bar[fctx][5], fpar);                      // the last MMA of the group: everything before it is complete
      fence_after_sync();
      {
        float v[TC_FW];
        load_sum2(tm_lane + TM_W1 + 8 + TC_FW * hf, v);
#pragma unroll
        for (int i = 0; i < TC_FW; ++i) acc2(g1[i], g1e[i], v[i] * inv_w1);
        uint32_t o4[4];
        tmem_ld4(tm_lane + TM_W1, o4);
        tmem_ld_wait();
        acc2(gb1, gb1e, __uint_as_float(o4[0]) * inv_b1);
      }
      {
        uint32_t a[TC_XW], b[TC_XW], o4[4];
        if constexpr (TC_XW == 8) {
          tmem_ld8(tm_lane + TM_W0 + 8 + 8 * hf, *reinterpret_cast<uint32_t(*)[8]>(a));
          tmem_ld8(tm_lane + TM_W0 + 8 + 16 + 8 * hf, *reinterpret_cast<uint32_t(*)[8]>(b));
        } else {
          tmem_ld4(tm_lane + TM_W0 + 8 + 4 * hf, *reinterpret_cast<uint32_t(*)[4]>(a));
          tmem_ld4(tm_lane + TM_W0 + 8 + 16 + 4 * hf, *reinterpret_cast<uint32_t(*)[4]>(b));
        }
        tmem_ld4(tm_lane + TM_W0, o4);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < TC_XW; ++i) acc2(g0[i], g0e[i], (__uint_as_float(b[i]) + __uint_as_float(a[i])) * inv_w0);
        acc2(gb0, gb0e, __uint_as_float(o4[0]) * inv_b0);
      }
      fence_before_sync();                                   // ordered before the next pair's MMAs by the phase_done arrivals
    fold_due = false;
  };
  long tile0 = blockIdx.x;                       // context c works on tile0 + c * gridDim.x
  prefetch_x(tile0, xa0, xb0, yn0);
  prefetch_x(tile0 + gridDim.x, xa1, xb1, yn1);
  uint32_t par = 0;
  int pair = 0;
  TC_STAMP(20);
  for (; tile0 < n_tiles; tile0 += 2L * gridDim.x, par ^= 1u, ++pair) {
    const int n_ctx = (tile0 + gridDim.x < n_tiles) ? 2 : 1;            // uniform over the CTA
    const bool fold = ((pair + 1) % TC_FLUSH == 0) || (tile0 + 2L * gridDim.x >= n_tiles);
    float yv0 = 0.f, yv1 = 0.f;
    // ---- P0: x tile -> fp16 pieces (rows beyond the data are zero) -----------------------------------------------------------
#pragma unroll 1
    for (int c = 0; c < n_ctx; ++c) {
      TcCtx& cx = s.ctx[c];
      if (pair > 0) mbar_wait(&s.bar[c][5], par ^ 1u);       // MMA5 of this context's previous tile has read xp and dl
      const float4 xa = c ? xa1 : xa0, xb = c ? xb1 : xb0;
      if (c) yv1 = yn1; else yv0 = yn0;
      if (hf < 2) {
        const float v[8] = {xa.x * s_x, xa.y * s_x, xa.z * s_x, xa.w * s_x, xb.x * s_x, xb.y * s_x, xb.z * s_x, xb.w * s_x};
        uint4 p1, p2;
        split2h(v, p1, p2);
        unsigned char* dst = reinterpret_cast<unsigned char*>(cx.xp) + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u + hf * TC_CS;
        *reinterpret_cast<uint4*>(dst) = p1;
        *reinterpret_cast<uint4*>(dst + TC_XP) = p2;
      }
      phase_done(&s.ready[c], lane);
    }
    prefetch_x(tile0 + 2L * gridDim.x, xa0, xb0, yn0);
    prefetch_x(tile0 + 3L * gridDim.x, xa1, xb1, yn1);
    TC_STAMP(1);
    // ---- P1: H1 = sigmoid(Z1 + b0) ------------------------------------------------------------------------------------
#pragma unroll 1
    for (int c = 0; c < n_ctx; ++c) {
      TcCtx& cx = s.ctx[c];
      mbar_wait(&s.bar[c][1], par);
      fence_after_sync();
      {
        float v[TC_FW];
        load_acc(tm_lane + TM_Z + 128 * c + TC_FW * hf, v);
#pragma unroll
        for (int j = 0; j < TC_FW; ++j)
          v[j] = TC_FMA_RCP(j) ? tc_sigmoid_fma<true>(fmaf(v[j], inv_z1, s.b0[TC_FW * hf + j]))
                         : tc_sigmoid<true>(fmaf(v[j], inv_z1, s.b0[TC_FW * hf + j]));     // H1 * TC_SH
        if constexpr (TC_FW == 32) {   // sh^2 H1 (1 - H1) parked in the free half of this context's Z columns until P3
          uint32_t m[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) m[j] = __float_as_uint((TC_SH - v[j]) * v[j]);
          tmem_st32(tm_lane + TM_Z + 128 * c + 64 + TC_FW * hf, m);
        }
        store_pieces32h(reinterpret_cast<unsigned char*>(cx.h1), (TC_FW / 8) * hf, r, v);
        if constexpr (TC_FW == 32) tmem_st_wait();
      }
      phase_done(&s.ready[c], lane);
    }
    TC_STAMP(3);
    if (fold_due) do_fold();
    // ---- P2: H2, head, log-likelihood, delta3, dW2, Delta2 -------------------------------------------------------------
#pragma unroll 1
    for (int c = 0; c < n_ctx; ++c) {
      TcCtx& cx = s.ctx[c];
      const long row0 = (tile0 + (long)c * gridDim.x) * DP_R;
      const int rows = (int)min((long)DP_R, n_rows - row0);
      const float yv = c ? yv1 : yv0;
      mbar_wait(&s.bar[c][2], par);
      fence_after_sync();
      float t[TC_FW], p_head = 0.5f, d_head = 0.f;
      {
        float h[TC_FW];
        load_acc(tm_lane + TM_Z + 128 * c + TC_FW * hf, h);
        float apart = 0.f;
#pragma unroll
        for (int j = 0; j < TC_FW; ++j) {
          h[j] = TC_FMA_RCP(j) ? tc_sigmoid_fma<false>(fmaf(h[j], inv_z2, s.b1[TC_FW * hf + j]))
                         : tc_sigmoid<false>(fmaf(h[j], inv_z2, s.b1[TC_FW * hf + j]));
          apart = fmaf(h[j], s.w2[TC_FW * hf + j], apart);
        }
        s.exch[c][hf][r] = apart;
        // everything of Delta2 that does not depend on the head runs before the rendez-vous with the other half of the row:
        // w2 sd H2 (1 - H2), with H2 (1 - H2) as one FMA (a single rounding of the exact value)
#pragma unroll
        for (int j = 0; j < TC_FW; ++j) t[j] = fmaf(-h[j], h[j], h[j]) * s.w2d[TC_FW * hf + j];
        asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "n"(32 * TC_CG) : "memory");   // the warps that share rows 32 q .. 32 q + 31
        float asum = s.exch[c][0][r];
#pragma unroll
        for (int cg = 1; cg < TC_CG; ++cg) asum += s.exch[c][cg][r];
        const float a = asum + s.b2;
        float d = 0.f;
        if (r < rows) {   // stats/loss.py:2 semantics, saturation -> NaN
          p_head = __fdividef(1.0f, 1.0f + __expf(-a));        // 2 ulp; exact 0 / 1 at saturation like the reference
          d = (p_head == 0.0f || p_head == 1.0f) ? NAN : (yv - p_head);
        }
        d_head = d;
#pragma unroll
        for (int j = 0; j < TC_FW; ++j) {
          const float dh = d * h[j];                                         // dW2 terms
          h[j] = d * t[j];                                                   // Delta2 * sd
          t[j] = dh;
        }
        store_pieces32h(reinterpret_cast<unsigned char*>(cx.dl), (TC_FW / 8) * hf, r, h);
      }
      phase_done(&s.ready[c], lane);
      // while the MMAs run: the log-likelihood term and the column sums for dW2
      if (hf == 0) {   // one branch-free log per row for hard labels; soft labels take the general form
        const float qv = (yv == 1.0f) ? p_head : 1.0f - p_head;
        float term = (yv == 1.0f && p_head == 1.0f) || (yv == 0.0f && p_head == 0.0f) ? NAN : __logf(qv);
        if (yv != 0.0f && yv != 1.0f) term = __logf(p_head) * yv + __logf(1.0f - p_head) * (1.0f - yv);
        if (r < rows) {
          acc2(ll, lle, term);
          acc2(gb2, gb2e, d_head);
        }
      }
      if constexpr (TC_FW == 32) {
        butterfly_step<16>(t, lane);
      } else {                                       // 16 values: fold the two half-warps first (both keep the sums)
#pragma unroll
        for (int i = 0; i < 16; ++i) t[i] += __shfl_xor_sync(0xffffffffu, t[i], 16);
      }
      butterfly_step<8>(t, lane);
      butterfly_step<4>(t, lane);
      butterfly_step<2>(t, lane);
      butterfly_step<1>(t, lane);
      acc2(gw2, gw2e, t[0]);                                                 // unit 32 hf + lane, rows of quadrant q
    }
    TC_STAMP(5);
    // ---- P3: Delta1 = D1 H1 (1 - H1), written over Delta2 once MMA4 has read it -------------------------------------------
#pragma unroll 1
    for (int c = 0; c < n_ctx; ++c) {
      TcCtx& cx = s.ctx[c];
      mbar_wait(&s.bar[c][3], par);
      fence_after_sync();
      {
        float v[TC_FW];
        if constexpr (TC_FW == 32) {   // D1 and the parked sh^2 H1 (1 - H1): two TMEM loads in flight together
          uint32_t d1[32], m[32];
          tmem_ld32(tm_lane + TM_Z + 128 * c + TC_FW * hf, d1);
          tmem_ld32(tm_lane + TM_Z + 128 * c + 64 + TC_FW * hf, m);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = (__uint_as_float(d1[j]) * c_d1) * __uint_as_float(m[j]);
        } else {
          load_acc(tm_lane + TM_Z + 128 * c + TC_FW * hf, v);
          // H1 of this row from its two fp16 pieces (exact up to 2^-22)
          const unsigned char* hp = reinterpret_cast<const unsigned char*>(cx.h1) + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u +
                                    (uint32_t)((TC_FW / 8) * hf) * TC_CS;
#pragma unroll
          for (int ch = 0; ch < TC_FW / 8; ++ch) {
            const uint4 p1 = *reinterpret_cast<const uint4*>(hp + ch * TC_CS);
            const uint4 p2 = *reinterpret_cast<const uint4*>(hp + ch * TC_CS + TC_ACT);
            const uint32_t w1[4] = {p1.x, p1.y, p1.z, p1.w}, w2v[4] = {p2.x, p2.y, p2.z, p2.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 hs = sum_f16x2(w1[k], w2v[k]);                     // H1 * TC_SH
              v[8 * ch + 2 * k] = (v[8 * ch + 2 * k] * c_d1) * ((TC_SH - hs.x) * hs.x);
              v[8 * ch + 2 * k + 1] = (v[8 * ch + 2 * k + 1] * c_d1) * ((TC_SH - hs.y) * hs.y);
            }
          }
        }
        mbar_wait(&s.bar[c][4], par);                        // MMA4 has read Delta2 (and H1)
        store_pieces32h(reinterpret_cast<unsigned char*>(cx.dl), (TC_FW / 8) * hf, r, v);
      }
      phase_done(&s.ready[c], lane);
    }
    TC_STAMP(9);
    // ---- P4: every TC_FLUSH pairs the weight-gradient sums are folded (do_fold, one phase late) -----------------------------------
    if (fold) { fold_due = true; fctx = n_ctx - 1; fpar = par; }
    TC_STAMP(12);
  }
  if (fold_due) do_fold();
  // every MMA has completed: the last fold waited for the last one issued

  // ---- this CTA's row of partial sums: staged in shared memory (the operand buffers are free now), written out coalesced ------
  fence_before_sync();
  epi_sync();
  static_assert(sizeof(TcCtx) >= sizeof(double) * (DP_P + 1) && sizeof(TcCtx) >= sizeof(double) * 3 * TC_THREADS, "staging space");
  double* row = reinterpret_cast<double*>(&s.ctx[0]);     // [DP_P + 1]
  double* red = reinterpret_cast<double*>(&s.ctx[1]);     // [3][TC_THREADS]
  if (lane < 16) {
    const int o = 16 * q + lane;
#pragma unroll
    for (int i = 0; i < TC_FW; ++i) row[1 + DP_OFF_W1 + o * DP_H + TC_FW * hf + i] = (double)g1[i] + (double)g1e[i];
#pragma unroll
    for (int j = 0; j < TC_XW; ++j) row[1 + o * DP_D0 + TC_XW * hf + j] = (double)g0[j] + (double)g0e[j];
    if (hf == 0) {
      row[1 + DP_OFF_B1 + o] = (double)gb1 + (double)gb1e;
      row[1 + DP_OFF_B0 + o] = (double)gb0 + (double)gb0e;
    }
  }
  // dW2: unit 32 hf + lane, partial over the rows of quadrant q; log-likelihood and db2: fixed-order block sums
  red[tid] = (double)gw2 + (double)gw2e;
  red[TC_THREADS + tid] = (double)ll + (double)lle;
  red[2 * TC_THREADS + tid] = (double)gb2 + (double)gb2e;
  epi_sync();
  if (tid < DP_H) {
    const int h2 = tid / TC_FW, l2 = tid % TC_FW;        // unit tid lives in column group h2, lane l2 of its warps
    double t = 0.0;
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) t += red[(4 * h2 + qq) * 32 + l2];
    row[1 + DP_OFF_W2 + tid] = t;
  }
  if (tid == 64) {
    double t = 0.0;
    for (int i = 0; i < DP_R; ++i) t += red[TC_THREADS + i];      // hf == 0 threads are tid 0..127
    row[0] = t;
  }
  if (tid == 96) {
    double t = 0.0;
    for (int i = 0; i < DP_R; ++i) t += red[2 * TC_THREADS + i];
    row[1 + DP_OFF_B2] = t;
  }
  epi_sync();
  for (int e = tid; e <= DP_P; e += TC_THREADS) out[e] = row[e];
  epi_sync();   // the next evaluation of a persistent run writes the operand buffers again
  TC_STAMP(21);
}

// The MMA-issue warp (first warp of the extra warp group; its three siblings only hold the group together for setmaxnreg).
// It walks through the same phases as the epilogue warps, waits until all eight have finished a phase of a context
// (s.ready[c]), and issues that phase's MMAs; completion goes back through tcgen05.commit -> s.bar[c][k].  Before round 2's
// last change warp 0 issued the MMAs behind a CTA barrier: ncu showed it spending a fifth of its time in UTCHMMA issue while
// the other seven warps waited for it at the next barrier (19 % of the kernel's warp-stall samples).
__device__ __forceinline__ void tc_issuer_body(TcSmem& s, const uint32_t tm, const long n_rows) {
  fence_async_smem();
  fence_before_sync();
  __syncthreads();           // pairs with the CTA barrier at the end of the epilogue warps' prologue
  fence_after_sync();
  if ((threadIdx.x >> 5) != TC_THREADS / 32) return;
  const long n_tiles = (n_rows + DP_R - 1) / DP_R;
  const uint32_t a_ctx0 = smem_u32(&s.ctx[0]);
  constexpr uint32_t CTXB = (uint32_t)sizeof(TcCtx);
  constexpr uint32_t OFF_H1 = (uint32_t)offsetof(TcCtx, h1), OFF_ONESH = (uint32_t)offsetof(TcCtx, ones_h);
  constexpr uint32_t OFF_XP = (uint32_t)offsetof(TcCtx, xp), OFF_ONESX = (uint32_t)offsetof(TcCtx, ones_x);
  constexpr uint32_t OFF_DL = (uint32_t)offsetof(TcCtx, dl);
  const uint32_t a_w0 = smem_u32(s.w0s), a_w1a = smem_u32(s.w1a), a_w1b = smem_u32(s.w1b);
  auto descs2 = [](uint32_t base, uint32_t piece_bytes, uint32_t lbo, uint32_t sbo, uint64_t (&d)[2]) {
    d[0] = smem_desc(base, lbo, sbo);
    d[1] = smem_desc(base + piece_bytes, lbo, sbo);
  };
  uint32_t rp0 = 0, rp1 = 0;                 // phase parities of s.ready[0 / 1]
  auto wait_ready = [&](int c) {
    uint32_t& rp = c ? rp1 : rp0;
    mbar_wait(&s.ready[c], rp);
    rp ^= 1u;
    fence_after_sync();
  };
  int pair = 0;
  for (long tile0 = blockIdx.x; tile0 < n_tiles; tile0 += 2L * gridDim.x, ++pair) {
    const int n_ctx = (tile0 + gridDim.x < n_tiles) ? 2 : 1;
#pragma unroll 1
    for (int c = 0; c < n_ctx; ++c) {        // after P0: MMA1  Z1 = X W0^T
      wait_ready(c);
      if (elect_one()) {
        uint64_t dA[2];
        descs2(a_ctx0 + c * CTXB + OFF_XP, TC_XP, TC_CS, 128, dA);                      // K-major A (M = row, K = feature)
        mma_product_sum<128, 64>(tm + TM_Z + 128 * c, dA, smem_desc(a_w0, TC_WCS2, 128), 0, 0, 1);
        mma_commit(&s.bar[c][1]);
      }
      __syncwarp();
    }
#pragma unroll 1
    for (int c = 0; c < n_ctx; ++c) {        // after P1: MMA2  Z2 = H1 W1^T
      wait_ready(c);
      if (elect_one()) {
        uint64_t dA[2];
        descs2(a_ctx0 + c * CTXB + OFF_H1, TC_ACT, TC_CS, 128, dA);                     // K-major A (M = row, K = hidden unit)
        mma_product_sum<128, 64>(tm + TM_Z + 128 * c, dA, smem_desc(a_w1a, TC_WCS2, 128), 2 * TC_CS, 2 * TC_WCS2, 4);
        mma_commit(&s.bar[c][2]);
      }
      __syncwarp();
    }
#pragma unroll 1
    for (int c = 0; c < n_ctx; ++c) {        // after P2: MMA3  D1 = Delta2 W1,  MMA4  [db1 dW1] = Delta2^T [1 H1]
      wait_ready(c);
      if (elect_one()) {
        const uint32_t keep = (c > 0 || pair % TC_FLUSH != 0) ? 1u : 0u;   // sums continue from the previous tile
        uint64_t dA[2], dM[2];
        descs2(a_ctx0 + c * CTXB + OFF_DL, TC_ACT, TC_CS, 128, dA);                     // K-major A (M = row, K = unit)
        descs2(a_ctx0 + c * CTXB + OFF_DL, TC_ACT, 128, TC_CS, dM);                     // MN-major A (M = unit, K = row)
        mma_product_sum<128, 64>(tm + TM_Z + 128 * c, dA, smem_desc(a_w1b, TC_WCS2, 128), 2 * TC_CS, 2 * TC_WCS2, 4);
        mma_commit(&s.bar[c][3]);
        mma_product<2, 0, 64, 8, 64, 1, 1>(tm + TM_W1, dM, smem_desc(a_ctx0 + c * CTXB + OFF_ONESH, 128, TC_CS), 256, 256, 8, keep);
        mma_commit(&s.bar[c][4]);
      }
      __syncwarp();
    }
#pragma unroll 1
    for (int c = 0; c < n_ctx; ++c) {        // after P3: MMA5  [db0 dW0] = Delta1^T [1 X]
      wait_ready(c);
      if (elect_one()) {
        const uint32_t keep = (c > 0 || pair % TC_FLUSH != 0) ? 1u : 0u;
        uint64_t dM[2];
        descs2(a_ctx0 + c * CTXB + OFF_DL, TC_ACT, 128, TC_CS, dM);                     // MN-major A (M = unit, K = row)
        mma_product<2, 0, 64, 8, 16, 1, 1>(tm + TM_W0, dM, smem_desc(a_ctx0 + c * CTXB + OFF_ONESX, 128, TC_CS), 256, 256, 8, keep);
        mma_commit(&s.bar[c][5]);
      }
      __syncwarp();
    }
  }
}

// partials: [gridDim.x][DP_P + 1] doubles
__global__ void __launch_bounds__(TC_LAUNCH, 1)
dp_eval_tc_kernel(const float* __restrict__ theta, const float* __restrict__ x, const float* __restrict__ y, long n_rows,
                  const float* __restrict__ x_absmax, double* __restrict__ partials) {
  extern __shared__ __align__(1024) unsigned char tc_raw[];
  TcSmem& s = *reinterpret_cast<TcSmem*>(tc_raw);
  const uint32_t tm = tc_setup(s);
  if (threadIdx.x >= TC_THREADS) {            // the issue warp group gives its registers to the epilogue warps
    reg_dec<TC_REG_ISSUE>();
    tc_issuer_body(s, tm, n_rows);
  } else {
    reg_inc<TC_REG_EPI>();
    tc_eval_body(s, tm, theta, x, y, n_rows, x_absmax, partials + (size_t)blockIdx.x * (DP_P + 1), false);
  }
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) tmem_dealloc<512>(tm);
}

// ---- the whole HMC run in ONE launch ---------------------------------------------------------------------------------------
// eeyore/samplers/hmc.py:100-170 (leapfrog, hamiltonian, linear-space accept) and the epoch loop of
// eeyore/samplers/serial_sampler.py:35-52 for the replicated chain of the data-sharded path.  Persistent CTAs, one per SM
// (cooperative launch); the trajectory never leaves the SMs:
//   per iteration:  momentum draw + first half step (post CTAs, thread = parameter)          -> grid barrier
//     L times:      evaluation over this CTA's row tiles (tc_eval_body) -> per-CTA partial sums -> grid barrier
//                   fold + peer-store exchange + prior + leapfrog (post CTAs, dp_post_entries) -> grid barrier
//                   (every CTA then re-reads theta' through L2 and re-splits the weights)
//                   accept test (every post CTA computes the same decision from the same scalars), commit, sample write-out
// The TMEM allocation, the ones blocks and the launch itself are paid once per run instead of once per evaluation, and there is
// no kernel boundary (launch gap) between an evaluation, its exchange and the next evaluation.
struct DpRunArgs {
  const float* x; const float* y; long n_rows; const float* x_absmax;
  float* theta_cur; float* grad_cur; double* target_cur;         // in/out chain state
  float* theta_p; float* grad_p; float* mom;                     // work vectors [P]
  double* partials;                                              // [gridDim.x][P + 1]
  double* scratch;                                               // DP_SCRATCH_LEN doubles (datapar_post.cuh)
  unsigned long long* grid_ctr;                                  // zeroed by the host before the launch
  int* status;
  const float* ploc; const float* pscale; int has_temp; double temp;
  float step; int num_steps;
  long n_iters, n_burnin;
  RngKey key; uint32_t iter0;
  const float* z_tape; const float* u_tape;                      // [n_iters][P], [n_iters] or NULL
  float* out_samples; double* out_target; uint8_t* out_acc;      // [n_saved][P], [n_saved], [n_saved] or NULL
  uint32_t* acc_count;
  DpExchange xc;                                                 // xc.seq = sequence number of the first evaluation
};

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// all CTAs of the (co-resident) grid; `epoch` counts the arrivals expected so far.  Called by the epilogue / post threads only.
__device__ __forceinline__ void grid_sync(unsigned long long* ctr, unsigned long long& epoch, int* status) {
  epi_sync();
  if (threadIdx.x == 0) {
    epoch += gridDim.x;
    __threadfence();
    atomicAdd(ctr, 1ull);
    long spins = 0;
    while (ld_acquire_gpu(ctr) < epoch) {
      __nanosleep(20);
      if (++spins > (1L << 27)) {     // seconds: a CTA is gone; fail loudly instead of hanging the box
        status[0] = 2;
        __trap();
      }
    }
  }
  epi_sync();
}

__global__ void __launch_bounds__(TC_LAUNCH, 1) dp_hmc_run_kernel(const DpRunArgs a) {
  extern __shared__ __align__(1024) unsigned char tc_raw[];
  TcSmem& s = *reinterpret_cast<TcSmem*>(tc_raw);
  __shared__ double red[DP_POST_THREADS / 32];
  __shared__ double bc[4];
  __shared__ int acc_s;
  static_assert(TC_THREADS == DP_POST_THREADS, "post layout: 256 threads per CTA");
  const int tid = threadIdx.x, cta = blockIdx.x;
  const bool post = cta < DP_POST_CTAS;                 // these CTAs also own the parameter entries
  const int e = cta * DP_POST_THREADS + tid, j = e - 1; // entry of [loglik, dloglik]; parameter j
  const bool own = post && e >= 1 && e <= DP_P;
  const uint32_t tm = tc_setup(s);
  if (tid >= TC_THREADS) {   // the issue warp group: one tc_issuer_body per evaluation, nothing else (no post work, no grid barrier)
    reg_dec<TC_REG_ISSUE>();
    for (long ev = 0; ev < a.n_iters * (long)a.num_steps; ++ev) tc_issuer_body(s, tm, a.n_rows);
    __syncthreads();
    return;
  }
  reg_inc<TC_REG_EPI>();
  unsigned long long epoch = 0;
  unsigned long long seq = a.xc.seq;
  double lt_cur = __ldcg(a.target_cur);
  bool again = false;
  uint32_t n_acc = 0;

  for (long it = 0; it < a.n_iters; ++it) {
    const uint32_t iter = a.iter0 + (uint32_t)it;
    // ---- momentum draw, kinetic energy, first half step: p = z + step/2 grad; theta' = theta + step p   (hmc.py:134,105,110)
    if (post) {
      double k = 0.0;
      if (own) {
        float z;
        if (a.z_tape) {
          z = a.z_tape[(size_t)it * DP_P + j];
        } else {   // the stream of dp_hmc_begin_kernel: four normals per Philox block
          U4 w = philox4x32_10(U4{(uint32_t)(j >> 2), iter, 0u, 0u}, a.key.k0, a.key.k1);
          float z0, z1;
          if ((j & 2) == 0) box_muller<float>(Uni<float>::from(w.x), Uni<float>::from(w.y), &z0, &z1);
          else box_muller<float>(Uni<float>::from(w.z), Uni<float>::from(w.w), &z0, &z1);
          z = (j & 1) ? z1 : z0;
        }
        k = (double)z * (double)z;
        const float pj = fmaf(0.5f * a.step, a.grad_cur[j], z);
        a.mom[j] = pj;
        a.theta_p[j] = fmaf(a.step, pj, a.theta_cur[j]);
      }
      const double kc = block_sum_256(k, red);
      if (tid == 0) a.scratch[DP_SCRATCH_KIN0 + 32 * (int)(it & 1) + cta] = kc;
    }
    grid_sync(a.grid_ctr, epoch, a.status);
    // ---- the trajectory ------------------------------------------------------------------------------------------------------
    for (int st = 0; st < a.num_steps; ++st) {
      tc_eval_body(s, tm, a.theta_p, a.x, a.y, a.n_rows, a.x_absmax, a.partials + (size_t)cta * (DP_P + 1), again);
      again = true;
      grid_sync(a.grid_ctr, epoch, a.status);
      if (post)
        dp_post_entries(a.partials, (int)gridDim.x, a.xc, seq, a.theta_p, a.ploc, a.pscale, a.has_temp, a.temp, a.grad_p,
                        st == a.num_steps - 1 ? 2 : 1, a.step, a.mom, a.theta_p, a.scratch, a.status, red, cta, tid);
      ++seq;
      grid_sync(a.grid_ctr, epoch, a.status);
    }
    // ---- accept test in linear space (hmc.py:143-148), commit, sample write-out -----------------------------------------------
    if (post) {
      if (tid == 0) {
        double lt_p, kin1, kin0 = 0.0;
        dp_post_scalars(a.scratch, a.has_temp, a.temp, lt_p, kin1);
        for (int c = 0; c < DP_POST_CTAS; ++c) kin0 += __ldcg(&a.scratch[DP_SCRATCH_KIN0 + 32 * (int)(it & 1) + c]);
        kin0 *= 0.5;
        const double h_cur = -lt_cur + kin0, h_prop = -lt_p + kin1;
        double rate = exp(h_cur - h_prop);
        rate = (rate > 1.0) ? 1.0 : rate;
        const float u = a.u_tape ? a.u_tape[it] : philox_uniform<float>(a.key, 0u, iter);
        acc_s = ((double)u < rate) ? 1 : 0;
        bc[0] = lt_p;
      }
      epi_sync();
      const int acc = acc_s;
      if (acc) lt_cur = bc[0];
      n_acc += (uint32_t)acc;
      const long k = it - a.n_burnin;
      if (own) {
        if (acc) { a.theta_cur[j] = a.theta_p[j]; a.grad_cur[j] = a.grad_p[j]; }
        if (k >= 0 && a.out_samples) a.out_samples[(size_t)k * DP_P + j] = a.theta_cur[j];
      }
      if (cta == 0 && tid == 0 && k >= 0) {
        if (a.out_target) a.out_target[k] = lt_cur;
        if (a.out_acc) a.out_acc[k] = (uint8_t)acc;
      }
      epi_sync();   // acc_s / bc are rewritten in the next iteration
    }
  }
  if (cta == 0 && tid == 0) {
    a.target_cur[0] = lt_cur;
    if (a.acc_count) a.acc_count[0] += n_acc;
  }
  __syncthreads();
  if ((tid >> 5) == 0) tmem_dealloc<512>(tm);
}

// out[0] = max |x[i]| (out zeroed by the caller): non-negative floats order like their bit patterns, NaN sorts above inf
__global__ void dp_absmax_kernel(const float* __restrict__ x, long n, float* __restrict__ out) {
  float m = 0.f;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float a = fabsf(x[i]);
    m = (a > m || a != a) ? a : m;
  }
  unsigned int b = __float_as_uint(m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned int*>(out), b);
}

// out[e] = sum over CTAs of partials[cta][e], fixed order
__global__ void dp_reduce_tc_kernel(const double* __restrict__ partials, int n_parts, double* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e > DP_P) return;
  double t = 0.0;
  for (int c = 0; c < n_parts; ++c) t += partials[(size_t)c * (DP_P + 1) + e];
  out[e] = t;
}

}  // namespace eb

using namespace eb;

extern "C" {

int eeyore_b200_set_error_(int code, const char* msg);

int eeyore_b200_dp_loglik_grad_x(const void* theta, const void* x, const void* y, int64_t n_rows, const void* x_absmax,
                                 void* out_sums, void* workspace, void* stream) {
  if (!theta || !x || !y || (!out_sums && !workspace) || n_rows < 1)
    return eeyore_b200_set_error_(EEYORE_B200_EINVAL, "dp_loglik_grad: bad argument");
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15))
    return eeyore_b200_set_error_(EEYORE_B200_EINVAL, "dp_loglik_grad: x and y must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  auto fail = [](cudaError_t e, const char* where) {
    std::string m = std::string(where) + ": " + cudaGetErrorString(e);
    return eeyore_b200_set_error_(EEYORE_B200_ECUDA, m.c_str());
  };
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long n_tiles = (n_rows + DP_R - 1) / DP_R;
  const int grid = (int)(n_tiles < sms ? n_tiles : sms);
  double* partials = (double*)workspace;     // caller-owned [SMs, P + 1] doubles, or NULL: stream-ordered temporary
  cudaError_t e;
  if (!workspace) {
    e = cudaMallocAsync((void**)&partials, sizeof(double) * (size_t)grid * (DP_P + 1), st);
    if (e != cudaSuccess) return fail(e, "dp_loglik_grad(alloc)");
  }
  float* absmax = (float*)x_absmax;          // caller-owned device scalar (max |x| of the shard), or NULL: computed here
  if (!x_absmax) {
    e = cudaMallocAsync((void**)&absmax, sizeof(float), st);
    if (e != cudaSuccess) return fail(e, "dp_loglik_grad(alloc)");
    cudaMemsetAsync(absmax, 0, sizeof(float), st);
    dp_absmax_kernel<<<sms * 4, 256, 0, st>>>((const float*)x, (long)n_rows * DP_D0, absmax);
  }
  static bool attr_set = false;
  if (!attr_set) {
    e = cudaFuncSetAttribute(dp_eval_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TcSmem));
    if (e != cudaSuccess) return fail(e, "dp_loglik_grad(attr)");
    attr_set = true;
  }
  dp_eval_tc_kernel<<<grid, TC_LAUNCH, sizeof(TcSmem), st>>>((const float*)theta, (const float*)x, (const float*)y,
                                                              (long)n_rows, absmax, partials);
  if (out_sums)   // NULL: the caller folds the per-CTA rows itself (dp_post does, together with the exchange step)
    dp_reduce_tc_kernel<<<(DP_P + 1 + 255) / 256, 256, 0, st>>>(partials, grid, (double*)out_sums);
  e = cudaGetLastError();
  if (!workspace) cudaFreeAsync(partials, st);
  if (!x_absmax) cudaFreeAsync(absmax, st);
  if (e != cudaSuccess) return fail(e, "dp_loglik_grad");
  return EEYORE_B200_OK;
}

int eeyore_b200_dp_hmc_run(const void* x, const void* y, int64_t n_rows, const void* x_absmax, void* theta_cur, void* grad_cur,
                           void* target_cur, void* theta_prop, void* grad_prop, void* momentum, void* workspace,
                           void* local_scratch, void* grid_counter, int32_t* status, const void* prior_loc,
                           const void* prior_scale, int has_temperature, double temperature, double step, int num_steps,
                           int64_t n_iters, int64_t n_burnin, uint64_t seed, uint64_t iter_offset, const void* z_tape,
                           const void* u_tape, void* out_samples, void* out_target, uint8_t* out_accepted,
                           uint32_t* accept_count, int world, int rank, uint64_t seq, void* const* peer_bases, void* stream) {
  if (!x || !y || n_rows < 1 || !x_absmax || !theta_cur || !grad_cur || !target_cur || !theta_prop || !grad_prop || !momentum ||
      !workspace || !local_scratch || !grid_counter || !status || !prior_loc || !prior_scale)
    return eeyore_b200_set_error_(EEYORE_B200_EINVAL, "dp_hmc_run: null argument");
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15))
    return eeyore_b200_set_error_(EEYORE_B200_EINVAL, "dp_hmc_run: x and y must be 16-byte aligned");
  if (num_steps < 1 || n_iters < 0 || !(step > 0)) return eeyore_b200_set_error_(EEYORE_B200_EINVAL, "dp_hmc_run: bad step / num_steps / n_iters");
  if (world < 1 || world > DP_XMAXW || rank < 0 || rank >= world || (world > 1 && (!peer_bases || seq == 0)))
    return eeyore_b200_set_error_(EEYORE_B200_EINVAL, "dp_hmc_run: bad world / rank / sequence number");
  if (iter_offset + (uint64_t)n_iters > (1ull << 32))
    return eeyore_b200_set_error_(EEYORE_B200_EINVAL, "dp_hmc_run: iter_offset + n_iters must stay below 2^32 (32-bit Philox counter words)");
  if (n_iters == 0) return EEYORE_B200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  auto fail = [](cudaError_t e, const char* where) {
    std::string m = std::string(where) + ": " + cudaGetErrorString(e);
    return eeyore_b200_set_error_(EEYORE_B200_ECUDA, m.c_str());
  };
  int dev = 0, sms = 0, coop = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  if (!coop || sms < DP_POST_CTAS) return eeyore_b200_set_error_(EEYORE_B200_EUNSUPPORTED, "dp_hmc_run: cooperative launch unavailable");
  static bool attr_set = false;
  cudaError_t e;
  if (!attr_set) {
    e = cudaFuncSetAttribute(dp_hmc_run_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TcSmem));
    if (e != cudaSuccess) return fail(e, "dp_hmc_run(attr)");
    attr_set = true;
  }
  DpRunArgs a{};
  a.x = (const float*)x; a.y = (const float*)y; a.n_rows = (long)n_rows; a.x_absmax = (const float*)x_absmax;
  a.theta_cur = (float*)theta_cur; a.grad_cur = (float*)grad_cur; a.target_cur = (double*)target_cur;
  a.theta_p = (float*)theta_prop; a.grad_p = (float*)grad_prop; a.mom = (float*)momentum;
  a.partials = (double*)workspace; a.scratch = (double*)local_scratch; a.grid_ctr = (unsigned long long*)grid_counter;
  a.status = status; a.ploc = (const float*)prior_loc; a.pscale = (const float*)prior_scale;
  a.has_temp = has_temperature; a.temp = temperature; a.step = (float)step; a.num_steps = num_steps;
  a.n_iters = (long)n_iters; a.n_burnin = (long)n_burnin;
  a.key = RngKey{(uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32)}; a.iter0 = (uint32_t)iter_offset;
  a.z_tape = (const float*)z_tape; a.u_tape = (const float*)u_tape;
  a.out_samples = (float*)out_samples; a.out_target = (double*)out_target; a.out_acc = out_accepted; a.acc_count = accept_count;
  a.xc.world = world; a.xc.rank = rank; a.xc.seq = seq;
  for (int p = 0; p < world && world > 1; ++p) {
    a.xc.inbox[p] = reinterpret_cast<double*>(peer_bases[p]);
    a.xc.flags[p] = reinterpret_cast<unsigned long long*>(a.xc.inbox[p] + 2 * DP_XMAXW * DP_XSLOT);
  }
  e = cudaMemsetAsync(grid_counter, 0, sizeof(unsigned long long), st);
  if (e != cudaSuccess) return fail(e, "dp_hmc_run(memset)");
  void* params[] = {(void*)&a};
  // every SM gets one CTA (195 KB of shared memory each): the grid barrier needs all of them resident, which the cooperative
  // launch guarantees (it fails instead of dead-locking when the device is shared)
  e = cudaLaunchCooperativeKernel((const void*)dp_hmc_run_kernel, dim3((unsigned)sms), dim3(TC_LAUNCH), params, sizeof(TcSmem), st);
  if (e != cudaSuccess) return fail(e, "dp_hmc_run");
  return EEYORE_B200_OK;
}

int eeyore_b200_dp_loglik_grad(const void* theta, const void* x, const void* y, int64_t n_rows, void* out_sums,
                               void* workspace, void* stream) {
  return eeyore_b200_dp_loglik_grad_x(theta, x, y, n_rows, nullptr, out_sums, workspace, stream);
}

int eeyore_b200_dp_absmax(const void* x, int64_t n_values, void* out_absmax, void* stream) {
  if (!x || !out_absmax || n_values < 1) return eeyore_b200_set_error_(EEYORE_B200_EINVAL, "dp_absmax: bad argument");
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaMemsetAsync(out_absmax, 0, sizeof(float), (cudaStream_t)stream);
  dp_absmax_kernel<<<sms * 4, 256, 0, (cudaStream_t)stream>>>((const float*)x, (long)n_values, (float*)out_absmax);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    std::string m = std::string("dp_absmax: ") + cudaGetErrorString(e);
    return eeyore_b200_set_error_(EEYORE_B200_ECUDA, m.c_str());
  }
  return EEYORE_B200_OK;
}

#ifdef DP_TC_PROFILE
/* debug builds only: cycles spent per phase by one warp of CTA 0, accumulated since the last call */
int eeyore_b200_dp_tc_profile(unsigned long long* out24) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out24, dp_tc_prof, sizeof(unsigned long long) * 24);
  unsigned long long zero[24] = {0};
  cudaMemcpyToSymbol(dp_tc_prof, zero, sizeof(zero));
  return 0;
}
#endif

}  // extern "C"
