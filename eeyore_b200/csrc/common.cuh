// Shared device/host helpers for the chain-batched kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <limits>
#include <string.h>

#if defined(__CUDACC__)
#define EB_HD __host__ __device__ __forceinline__
#define EB_D __device__ __forceinline__
#else
#define EB_HD inline
#define EB_D inline
#endif

namespace eb {

template <typename T> EB_HD T qnan() { return (T)NAN; }

template <typename T> EB_HD T exp_t(T v);
template <> EB_HD float exp_t<float>(float v) { return expf(v); }
template <> EB_HD double exp_t<double>(double v) { return exp(v); }
template <typename T> EB_HD T log_t(T v);
template <> EB_HD float log_t<float>(float v) { return logf(v); }
template <> EB_HD double log_t<double>(double v) { return log(v); }
template <typename T> EB_HD T sqrt_t(T v);
template <> EB_HD float sqrt_t<float>(float v) { return sqrtf(v); }
template <> EB_HD double sqrt_t<double>(double v) { return sqrt(v); }
template <typename T> EB_HD T fma_t(T a, T b, T c);
template <> EB_HD float fma_t<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <> EB_HD double fma_t<double>(double a, double b, double c) { return fma(a, b, c); }

// ---- fp64 fast paths ----------------------------------------------------------------------------------------------
// The chain kernels are bound by the FP64 pipe, and most FP64 instructions of an evaluation are spent inside exp() and
// the division of the sigmoid.  These versions keep ~1 ulp accuracy (parity tolerance is 1e-10) but
//   * use a 64-entry 2^(j/64) table in shared memory + a degree-5 polynomial (the CUDA math library evaluates a
//     degree-11 polynomial and materialises each 64-bit coefficient with two UMOVs per use, ~26% of all issued
//     instructions in the first version of the kernels),
//   * have no slow-path branches (arguments are clamped instead),
//   * refine MUFU.RCP64H with one cubic step instead of the IEEE-exact division sequence.
EB_HD int dbl_lo(double v) {
#if defined(__CUDA_ARCH__)
  return __double2loint(v);
#else
  uint64_t b; memcpy(&b, &v, 8); return (int)(uint32_t)(b & 0xffffffffu);
#endif
}
EB_HD double dbl_add_exponent(double v, int k) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(__double2hiint(v) + (k << 20), __double2loint(v));
#else
  uint64_t b; memcpy(&b, &v, 8); b += (uint64_t)((int64_t)k << 52); memcpy(&v, &b, 8); return v;
#endif
}

// Table of 2^(j/64).  Device: a per-CTA shared-memory copy filled by exp_table_init() in the kernel prologue (the lookup
// index differs per lane, which the constant bank would serialise); host (tests/hostsim): the plain array.
static const double kExp2TabHost[64] = {
#include "exp_table.inc"
};
#if defined(__CUDACC__)
static __constant__ double kExp2TabDev[64] = {
#include "exp_table.inc"
};
EB_D double* exp_table_smem() {
  __shared__ double tab[64];
  return tab;
}
// every thread of the CTA calls this once before the first fp64 sigmoid / softmax; followed by a __syncthreads()
EB_D void exp_table_init() {
  double* t = exp_table_smem();
  for (int j = threadIdx.x; j < 64; j += blockDim.x) t[j] = kExp2TabDev[j];
}
#endif

// exp(a) for |a| <= 700 (callers clamp): a = (64 k + j) ln2/64 + r, |r| <= ln2/128; exp(a) = 2^k * 2^(j/64) * P5(r).
// 10 FP64-pipe instructions (the degree-11 single-polynomial version needed 15); ~1.5 ulp; NaN propagates.
// constants with non-zero low words would otherwise be materialised by UMOV pairs at every use
#define EB_MATH_CONSTS                                                                                             \
  {92.332482616893657, -1.08304246932675596e-02, -2.98158582698529328e-12, 8.3333333333333332e-03,                \
   4.1666666666666664e-02, 1.6666666666666666e-01,                                                                \
   /* log: ln2_hi, ln2_lo, Lg1..Lg7 (Sun fdlibm e_log.c) */                                                       \
   6.93147180369123816490e-01, 1.90821492927058770002e-10, 6.666666666666735130e-01, 3.999999999940941908e-01,    \
   2.857142874366239149e-01, 2.222219843214978396e-01, 1.818357216161805012e-01, 1.531383769920937332e-01,        \
   1.479819860511658591e-01}
static const double kMathHost[15] = EB_MATH_CONSTS;
#if defined(__CUDACC__)
static __constant__ double kMathDev[15] = EB_MATH_CONSTS;
#endif
EB_HD const double* math_consts() {
#if defined(__CUDA_ARCH__)
  return kMathDev;
#else
  return kMathHost;
#endif
}

EB_HD double exp_core(double a) {
  const double* c = math_consts();
  const double magic = 6755399441055744.0;  // 1.5 * 2^52: adding it rounds to the nearest integer
  const double t = fma(a, c[0], magic);                            // 64 / ln 2
  const double nf = t - magic;
  const int n = dbl_lo(t);
  double r = fma(nf, c[1], a);                                     // -ln2_hi / 64 (32 significant bits: nf * hi is exact)
  r = fma(nf, c[2], r);                                            // -ln2_lo / 64
#if defined(__CUDA_ARCH__)
  const double tj = exp_table_smem()[n & 63];
#else
  const double tj = kExp2TabHost[n & 63];
#endif
  double p = fma(r, c[3], c[4]);
  p = fma(p, r, c[5]);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return dbl_add_exponent(tj * p, n >> 6);
}

// exp(a) for a <= 0 (softmax numerators): anything below e^-700 is far under one ulp of the sum it is added to.
EB_HD double exp_nonpos(double a) { return exp_core(a < -700.0 ? -700.0 : a); }

template <typename T> EB_HD T exp_nonpos_t(T a) { return exp_t<T>(a); }
template <> EB_HD double exp_nonpos_t<double>(double a) { return exp_nonpos(a); }

// 1/d for finite d >= 1: MUFU.RCP64H seed + one cubic (Halley-type) refinement, 3 DFMA.
EB_HD double rcp_ge1(double d) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  // one third-order step: r <- r (1 + e + e^2), e = 1 - d r; seed error < 2^-20 -> < 2^-60
  const double e = fma(-d, r, 1.0);
  r = fma(r, fma(e, e, e), r);
  return r;
#else
  return 1.0 / d;
#endif
}

// log(x) for positive, normal, finite x (probabilities in (0, 1]); branch-free restatement of the classic
// fdlibm algorithm: x = 2^k (1 + f), sqrt(1/2) <= 1 + f < sqrt(2); s = f / (2 + f);
// log(1 + f) = f - f^2/2 + s (f^2/2 + R(s^2)).  < 1 ulp.  Callers handle 0 / NaN.
EB_HD double log_pos_normal(double x) {
  const double* c = math_consts();
#if defined(__CUDA_ARCH__)
  int hx = __double2hiint(x);
  const int lx = __double2loint(x);
#else
  uint64_t b; memcpy(&b, &x, 8);
  int hx = (int)(b >> 32);
  const int lx = (int)(uint32_t)(b & 0xffffffffu);
#endif
  int k = (hx >> 20) - 1023;
  hx &= 0x000fffff;
  const int i = (hx + 0x95f64) & 0x100000;
  k += i >> 20;
#if defined(__CUDA_ARCH__)
  const double m = __hiloint2double(hx | (i ^ 0x3ff00000), lx);
#else
  b = ((uint64_t)(uint32_t)(hx | (i ^ 0x3ff00000)) << 32) | (uint32_t)lx;
  double m; memcpy(&m, &b, 8);
#endif
  const double f = m - 1.0;
  const double s = f * rcp_ge1(2.0 + f);
  const double dk = (double)k;
  const double z = s * s;
  const double w = z * z;
  const double t1 = w * fma(w, fma(w, c[13], c[11]), c[9]);
  const double t2 = z * fma(w, fma(w, fma(w, c[14], c[12]), c[10]), c[8]);
  const double R = t2 + t1;
  const double hfsq = 0.5 * f * f;
  return fma(dk, c[6], -((hfsq - fma(s, hfsq + R, dk * c[7])) - f));
}

// sigmoid as the reference evaluates it, 1 / (1 + exp(-g))  (torch.sigmoid, eeyore/models/mlp.py:48-49).
template <typename T> EB_HD T sigmoid_t(T g) { return T(1) / (T(1) + exp_t<T>(-g)); }
// fp64: exp(-g) is clamped to [e^-700, e^700]; below, 1 + e == 1 exactly as in the reference; above, the result is
// ~1e-304 where the reference underflows towards 0 -- the head of the network restores the exact-zero case (head_loss).
template <> EB_HD double sigmoid_t<double>(double g) {
  double a = -g;
  a = (fabs(a) > 700.0) ? copysign(700.0, a) : a;  // NaN compares false and propagates
  return rcp_ge1(1.0 + exp_core(a));
}

// N sigmoids evaluated stage by stage ("vertically"): the N dependency chains are written interleaved so that the
// instruction scheduler keeps all of them in flight (a hidden layer's units are independent; each chain alone is
// latency-bound: ~14 dependent FP64 operations plus a table lookup and a MUFU).
template <typename T, int N> EB_HD void sigmoid_vec(const T (&g)[N], T (&out)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) out[i] = sigmoid_t<T>(g[i]);
}
#if defined(__CUDA_ARCH__)
template <int N> EB_D void sigmoid_vec_f64(const double (&g)[N], double (&out)[N]) {
  const double* c = math_consts();
  const double magic = 6755399441055744.0;
  double a[N], t[N], nf[N], r[N], p[N], tj[N], d[N], q[N], e[N];
  int n[N];
#pragma unroll
  for (int i = 0; i < N; ++i) { a[i] = -g[i]; a[i] = (fabs(a[i]) > 700.0) ? copysign(700.0, a[i]) : a[i]; }
#pragma unroll
  for (int i = 0; i < N; ++i) t[i] = fma(a[i], c[0], magic);
#pragma unroll
  for (int i = 0; i < N; ++i) { nf[i] = t[i] - magic; n[i] = __double2loint(t[i]); }
#pragma unroll
  for (int i = 0; i < N; ++i) tj[i] = exp_table_smem()[n[i] & 63];
#pragma unroll
  for (int i = 0; i < N; ++i) r[i] = fma(nf[i], c[1], a[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) r[i] = fma(nf[i], c[2], r[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) p[i] = fma(r[i], c[3], c[4]);
#pragma unroll
  for (int i = 0; i < N; ++i) p[i] = fma(p[i], r[i], c[5]);
#pragma unroll
  for (int i = 0; i < N; ++i) p[i] = fma(p[i], r[i], 0.5);
#pragma unroll
  for (int i = 0; i < N; ++i) p[i] = fma(p[i], r[i], 1.0);
#pragma unroll
  for (int i = 0; i < N; ++i) p[i] = fma(p[i], r[i], 1.0);
#pragma unroll
  for (int i = 0; i < N; ++i) d[i] = 1.0 + dbl_add_exponent(tj[i] * p[i], n[i] >> 6);
#pragma unroll
  for (int i = 0; i < N; ++i) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(q[i]) : "d"(d[i]));
#pragma unroll
  for (int i = 0; i < N; ++i) e[i] = fma(-d[i], q[i], 1.0);
#pragma unroll
  for (int i = 0; i < N; ++i) e[i] = fma(e[i], e[i], e[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) out[i] = fma(q[i], e[i], q[i]);
}
#endif


// cos/sin(2 pi u)
template <typename T> EB_HD void sincos2pi(T u, T* s, T* c);
template <> EB_HD void sincos2pi<float>(float u, float* s, float* c) {
#if defined(__CUDA_ARCH__)
  sincospif(2.0f * u, s, c);
#else
  *s = sinf(6.283185307179586f * u); *c = cosf(6.283185307179586f * u);
#endif
}
template <> EB_HD void sincos2pi<double>(double u, double* s, double* c) {
#if defined(__CUDA_ARCH__)
  sincospi(2.0 * u, s, c);
#else
  *s = sin(6.283185307179586 * u); *c = cos(6.283185307179586 * u);
#endif
}

#if defined(__CUDACC__)
// xor-butterfly all-reduce over the G lanes of a chain group (G a power of two <= 32).
// fp add is commutative, so every lane ends with the bitwise-identical sum.
// Lane mask of the G-lane chain group the calling lane belongs to.  Groups of one warp may diverge from each other
// (per-chain leapfrog counts under the dual-averaging tuner), so every intra-group primitive names only its own lanes.
template <int G> EB_D unsigned group_mask() {
  if (G >= 32) return 0xffffffffu;
  return ((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1));
}
template <int G, typename T> EB_D T group_allreduce(T v) {
  const unsigned mask = group_mask<G>();
#pragma unroll
  for (int m = G / 2; m >= 1; m >>= 1) v += __shfl_xor_sync(mask, v, m);
  return v;
}
#endif

}  // namespace eb
