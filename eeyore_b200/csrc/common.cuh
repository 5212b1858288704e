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
// The chain kernels are bound by the FP64 pipe (a warp-wide FP64 instruction occupies it for two cycles), and most FP64
// instructions of an evaluation used to be spent inside exp(), the division of the sigmoid and log().  These versions keep
// a few ulp accuracy (parity tolerance is 1e-10) with as few FP64 instructions and as short a dependency chain as possible:
//   * exp: 2048-entry 2^(j/2048) table in shared memory + a degree-3 polynomial; the argument reduction is one exact
//     FMA in base 2 (r = a * 2048/ln2 - n; the rounding of the constant costs |a| * 2^-53 relative);
//   * sigmoid: 1 + e is formed by one FMA from the exponent-adjusted table entry, MUFU.RCP64H is refined with one cubic
//     step, saturation is an integer test off the critical path: 10 FP64 instructions (CUDA libm + IEEE division: ~45);
//   * log: fdlibm-style normalisation to [sqrt(1/2), sqrt(2)), a 1024-entry (1/c, -log(1/c)) table and a degree-4
//     polynomial, no division: 7 FP64 instructions (fdlibm: 23 and a reciprocal).
// Tables: tools/gen_math_tables.py (correctly rounded with mpmath).
EB_HD int dbl_lo(double v) {
#if defined(__CUDA_ARCH__)
  return __double2loint(v);
#else
  uint64_t b; memcpy(&b, &v, 8); return (int)(uint32_t)(b & 0xffffffffu);
#endif
}
EB_HD int dbl_hi(double v) {
#if defined(__CUDA_ARCH__)
  return __double2hiint(v);
#else
  uint64_t b; memcpy(&b, &v, 8); return (int)(uint32_t)(b >> 32);
#endif
}
EB_HD double dbl_from(int hi, int lo) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(hi, lo);
#else
  uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double v; memcpy(&v, &b, 8); return v;
#endif
}
EB_HD double dbl_add_exponent(double v, int k) { return dbl_from(dbl_hi(v) + (int)((unsigned)k << 20), dbl_lo(v)); }

constexpr int kExpTabN = 2048, kLogTabN = 1024;
// Host (tests/hostsim): the plain arrays.  Device: global-memory masters, copied by every CTA into shared memory in its
// prologue (the lookup index differs per lane, which the constant bank would serialise).
static const double kExp2TabHost[kExpTabN] = {
#include "exp_table.inc"
};
#include "log_table.inc"
static const double kLogInvHost[kLogTabN] = {EB_LOG_INV_TABLE};
static const double kLogNlogHost[kLogTabN] = {EB_LOG_NLOG_TABLE};
#if defined(__CUDACC__)
static __device__ const double kExp2TabDev[kExpTabN] = {
#include "exp_table.inc"
};
static __device__ const double kLogInvDev[kLogTabN] = {EB_LOG_INV_TABLE};
static __device__ const double kLogNlogDev[kLogTabN] = {EB_LOG_NLOG_TABLE};
// [2^(j/2048) (2048) | 1/c (1024) | -log(1/c) (1024)]: 32 KB of static shared memory in every kernel that evaluates fp64
// sigmoids / logs
EB_D double* exp_table_smem() {
  __shared__ double tab[kExpTabN + 2 * kLogTabN];
  return tab;
}
// every thread of the CTA calls this once before the first fp64 sigmoid / softmax / log; followed by a __syncthreads()
EB_D void exp_table_init() {
  double* t = exp_table_smem();
  for (int j = threadIdx.x; j < kExpTabN; j += blockDim.x) t[j] = kExp2TabDev[j];
  for (int j = threadIdx.x; j < kLogTabN; j += blockDim.x) {
    t[kExpTabN + j] = kLogInvDev[j];
    t[kExpTabN + kLogTabN + j] = kLogNlogDev[j];
  }
}
#endif
EB_HD double exp2_tab(int j) {
#if defined(__CUDA_ARCH__)
  return exp_table_smem()[j];
#else
  return kExp2TabHost[j];
#endif
}
EB_HD double log_inv_tab(int i) {
#if defined(__CUDA_ARCH__)
  return exp_table_smem()[kExpTabN + i];
#else
  return kLogInvHost[i];
#endif
}
EB_HD double log_nlog_tab(int i) {
#if defined(__CUDA_ARCH__)
  return exp_table_smem()[kExpTabN + kLogTabN + i];
#else
  return kLogNlogHost[i];
#endif
}

// constants with non-zero low words live in the constant bank (DFMA takes c[][] operands; immediates would be
// materialised by UMOV pairs at every use): 2048/ln2; a1, a2, a3 of 2^(r/2048) = 1 + a1 r + a2 r^2 + a3 r^3; ln2; 1/3
#define EB_MATH_CONSTS                                                                                     \
  {2954.639443740597, 0.0003384507717577858, 5.727446245172041e-08, 6.461528672932366e-12, 0.6931471805599453, \
   0.3333333333333333,                                                                                     \
   /* [6..13] sin(pi r) = r (S0 + S1 z + ... + S7 z^7), [14..21] cos(pi r) = 1 + z (C1 + ... + C8 z^7), z = r^2 */ \
   3.141592653589793, -5.16771278004997, 2.5501640398773455, -0.5992645293207921, 0.08214588661112823,     \
   -0.0073704309457143504, 0.00046630280576761255, -2.1915353447830217e-05,                                \
   -4.934802200544679, 4.0587121264167685, -1.3352627688545895, 0.2353306303588932, -0.02580689139001406,  \
   0.0019295743094039231, -0.0001046381049248457, 4.303069587032947e-06}
static const double kMathHost[22] = EB_MATH_CONSTS;
#if defined(__CUDACC__)
static __constant__ double kMathDev[22] = EB_MATH_CONSTS;
#endif
EB_HD const double* math_consts() {
#if defined(__CUDA_ARCH__)
  return kMathDev;
#else
  return kMathHost;
#endif
}

// 2^k 2^(j/2048) for n = 2048 k + j: the table entry with k added to its exponent field.
EB_HD double exp2_tab_scaled(int n) {
#if defined(__CUDA_ARCH__)
  // five integer instructions spelled out (the compiler's own sequence for the same expression is seven; the dispatch
  // port is as scarce as the FP64 pipe here)
  int lo, hi;
  asm("{\n\t.reg .u32 a;\n\t.reg .s32 k;\n\t"
      "and.b32 a, %2, 2047;\n\t"
      "mad.lo.u32 a, a, 8, %3;\n\t"
      "ld.shared.v2.b32 {%0, %1}, [a];\n\t"
      "shr.s32 k, %2, 11;\n\t"
      "mad.lo.s32 %1, k, 1048576, %1;\n\t}"
      : "=r"(lo), "=r"(hi)
      : "r"(n), "r"((unsigned)__cvta_generic_to_shared(exp_table_smem())));
  return __hiloint2double(hi, lo);
#else
  return dbl_add_exponent(exp2_tab(n & (kExpTabN - 1)), n >> 11);
#endif
}

// The pieces of exp(a), |a| <= 708: a 2048/ln2 = n + r with |r| <= 1/2 (one exact FMA), n = 2048 k + j;
// exp(a) = [2^k 2^(j/2048)] * q,  q = 2^(r/2048) = 1 + r (a1 + r (a2 + r a3))  (truncation 3.4e-17).  5 FP64 instructions.
EB_HD void exp_pieces(double a, double& tjs, double& q) {
  const double* c = math_consts();
  const double magic = 6755399441055744.0;  // 1.5 * 2^52: adding it rounds to the nearest integer
  const double t = fma(a, c[0], magic);
  const double nf = t - magic;
  const int n = dbl_lo(t);
  const double r = fma(a, c[0], -nf);
  double p = fma(r, c[3], c[2]);
  p = fma(p, r, c[1]);
  q = fma(r, p, 1.0);
  tjs = exp2_tab_scaled(n);
}
EB_HD double exp_core(double a) {
  double tjs, q;
  exp_pieces(a, tjs, q);
#if defined(__CUDA_ARCH__)
  return __dmul_rn(tjs, q);   // never contracted into a consumer's add: GRAD / no-GRAD instantiations stay bit-identical
#else
  return tjs * q;
#endif
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

// log(x) for positive, normal, finite x (probabilities in (0, 1], uniforms, softmax sums).  x = 2^k m with
// sqrt(1/2) <= m < sqrt(2) (the high-word trick of fdlibm's e_log.c); the window position of m selects c_i with
// |m / c_i - 1| < 4.9e-4 (the bin around 1 has c = 1, so there is no cancellation for x -> 1);
// log x = k ln2 - log(1/c_i) + log1p(r),  r = m (1/c_i) - 1 (one FMA),  log1p(r) = r - r^2/2 + r^3/3 - r^4/4  (5.5e-18).
// Callers handle 0 / NaN.
EB_HD double log_pos_normal(double x) {
  const double* c = math_consts();
  const int hs = dbl_hi(x) + 0x95f64;
  const int f = hs & 0xfffff;
  const double m = dbl_from(f + (0x3ff00000 - 0x95f64), dbl_lo(x));
  const int i = f >> 10;
  const double dk = (double)((hs >> 20) - 1023);
  const double r = fma(m, log_inv_tab(i), -1.0);
  const double base = fma(dk, c[4], log_nlog_tab(i));
  const double s = r * r;
  double q = fma(r, -0.25, c[5]);
  q = fma(q, r, -0.5);
  return base + fma(s, q, r);
}

// The far negative tail of the head's sigmoid, evaluated like the reference (libm exp, IEEE division): for a below -708 the
// probability is subnormal (the fast reciprocal flushes it to 0) and exactly 0 from -709.78 on, where exp(-a) overflows.
// Only the general (rare) path of the head calls this; its log must then accept subnormal arguments as well.
EB_HD double sigmoid_ref_tail(double a) { return 1.0 / (1.0 + exp(-a)); }
EB_HD double log_prob_any(double q) { return dbl_hi(q) < 0x00100000 ? log(q) : log_pos_normal(q); }   // q > 0

// Exact tests on values that are known to be non-negative or NaN (probabilities): fp64 versions are integer compares
// on the two words -- a DSETP would occupy the FP64 pipe like a DFMA.
template <typename T> EB_HD bool prob_is_zero(T p) { return p == T(0); }
template <typename T> EB_HD bool prob_is_one(T p) { return p == T(1); }
template <typename T> EB_HD bool prob_is_nan(T p) { return p != p; }
template <> EB_HD bool prob_is_zero<double>(double p) { return (dbl_hi(p) | dbl_lo(p)) == 0; }
template <> EB_HD bool prob_is_one<double>(double p) { return dbl_hi(p) == 0x3ff00000 && dbl_lo(p) == 0; }
template <> EB_HD bool prob_is_nan<double>(double p) { return (dbl_hi(p) & 0x7ff00000) == 0x7ff00000; }  // never inf

// sigmoid as the reference evaluates it, 1 / (1 + exp(-g))  (torch.sigmoid, eeyore/models/mlp.py:48-49).
template <typename T> EB_HD T sigmoid_t(T g) { return T(1) / (T(1) + exp_t<T>(-g)); }
// fp64: d = 1 + exp(-g) by one FMA from the pieces of the exponential, then the refined reciprocal.  |g| > 708 (where the
// exponent arithmetic would wrap) is replaced at the end by the saturated value -- 1 for g > 0 (1 + e == 1 exactly from
// g > 36.7 on, as in the reference), 0 for g < -708 (the reference: below 3.3e-308, exactly 0 from -709.78 on); the test
// is integer arithmetic on the high word, off the dependency chain, false for NaN (which propagates through the FMAs).
template <> EB_HD double sigmoid_t<double>(double g) {
  double tjs, q;
  exp_pieces(-g, tjs, q);
  const double s = rcp_ge1(fma(tjs, q, 1.0));
  const int hi = dbl_hi(g);
  const bool big = (int)((unsigned)(hi & 0x7fffffff) + 0xfffffu) > (0x40862000 + 0xfffff);   // signed: NaN high words wrap negative
  const double sat = dbl_from(hi < 0 ? 0 : 0x3ff00000, 0);
  return big ? sat : s;
}

// The same without the saturation select: valid for |g| <= 708 and not NaN.  Callers (accumulate_row's fast path) track the
// largest high word of the arguments and redo the row with sigmoid_t when the bound is violated.
EB_HD double sigmoid_fast(double g) {
  double tjs, q;
  exp_pieces(-g, tjs, q);
  return rcp_ge1(fma(tjs, q, 1.0));
}
// high word of |g|: as an unsigned number it orders finite values, infinity and NaN like the magnitude
EB_HD int abs_hi(double g) { return dbl_hi(g) & 0x7fffffff; }
constexpr int kAbsHi708 = 0x40862000;   // high word of 708.0
constexpr int kAbsHi36 = 0x40420000;    // high word of 36.0: below, 0 < sigmoid < 1 strictly (1 + e^-36 > 1 in fp64)

// M independent fast sigmoids written stage by stage, so that the instruction scheduler sees M interleaved dependency
// chains (each alone is latency-bound: ten dependent FP64 operations, a table fetch and a MUFU).  Same arithmetic as
// sigmoid_fast; `mx` collects the largest |g| high word.
template <int M> EB_HD void sigmoid_fast_vec(const double (&g)[M], double (&out)[M], int& mx) {
  const double* c = math_consts();
  const double magic = 6755399441055744.0;
  double t[M], r[M], p[M], tj[M];
#pragma unroll
  for (int i = 0; i < M; ++i) t[i] = fma(-g[i], c[0], magic);
#pragma unroll
  for (int i = 0; i < M; ++i) tj[i] = exp2_tab_scaled(dbl_lo(t[i]));
#pragma unroll
  for (int i = 0; i < M; ++i) r[i] = fma(-g[i], c[0], -(t[i] - magic));
#pragma unroll
  for (int i = 0; i < M; ++i) p[i] = fma(r[i], c[3], c[2]);
#pragma unroll
  for (int i = 0; i < M; ++i) p[i] = fma(p[i], r[i], c[1]);
#pragma unroll
  for (int i = 0; i < M; ++i) p[i] = fma(r[i], p[i], 1.0);
#pragma unroll
  for (int i = 0; i < M; ++i) p[i] = fma(tj[i], p[i], 1.0);
#pragma unroll
  for (int i = 0; i < M; ++i) out[i] = rcp_ge1(p[i]);
#pragma unroll
  for (int i = 0; i < M; ++i) { const int ah = abs_hi(g[i]); mx = ah > mx ? ah : mx; }
}

// ---- fp32 fast paths ----------------------------------------------------------------------------------------------------
// The fp32 chain kernels are bound by instruction issue (one warp instruction per cycle and scheduler) and by the MUFU unit,
// not by an arithmetic pipe: libm's expf + IEEE division + logf cost ~45 instructions per sigmoid.  On the fast path (the head's
// pre-activation within +-16, where 0 < p < 1 strictly in fp32 and nothing saturates) a sigmoid is ex2.approx + add +
// rcp.approx (2 ulp each: 3e-7 relative, against a 1e-5 bar), a log is lg2.approx * ln 2; anything else goes through the
// general code, which reproduces the reference's saturation / NaN semantics with libm.
EB_HD float sigmoid_fast(float g) {
#if defined(__CUDA_ARCH__)
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(g * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
#else
  return 1.0f / (1.0f + expf(-g));
#endif
}
EB_HD float exp_nonpos_fast(float a) {
#if defined(__CUDA_ARCH__)
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a * 1.4426950408889634f));
  return e;
#else
  return expf(a);
#endif
}
EB_HD float rcp_fast(float s) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s));
  return r;
#else
  return 1.0f / s;
#endif
}
EB_HD float log_fast(float x) {
#if defined(__CUDA_ARCH__)
  return __logf(x);
#else
  return logf(x);
#endif
}
EB_HD int abs_hi(float g) {
#if defined(__CUDA_ARCH__)
  return __float_as_int(g) & 0x7fffffff;
#else
  uint32_t b; memcpy(&b, &g, 4); return (int)(b & 0x7fffffffu);
#endif
}
constexpr int kAbsBits16f = 0x41800000;   // bit pattern of 16.0f: below, 0 < sigmoid < 1 strictly in fp32
template <int M> EB_HD void sigmoid_fast_vec(const float (&g)[M], float (&out)[M], int& mx) {
#pragma unroll
  for (int i = 0; i < M; ++i) out[i] = sigmoid_fast(g[i]);
#pragma unroll
  for (int i = 0; i < M; ++i) { const int ah = abs_hi(g[i]); mx = ah > mx ? ah : mx; }
}
// bounds of the fast path per type: hidden pre-activations (fp64: the table-based exp wraps beyond 708; fp32: ex2.approx
// saturates correctly everywhere) and the head's (0 < p < 1 strictly)
template <typename T> struct FastBounds;
template <> struct FastBounds<double> { static constexpr int hidden = kAbsHi708, head = kAbsHi36; };
template <> struct FastBounds<float> { static constexpr int hidden = 0x7fffffff, head = kAbsBits16f; };
EB_HD double log_prob_fast(double q) { return log_pos_normal(q); }
EB_HD float log_prob_fast(float q) { return log_fast(q); }

// N sigmoids evaluated stage by stage ("vertically"): the N dependency chains are written interleaved so that the
// instruction scheduler keeps all of them in flight (a hidden layer's units are independent; each chain alone is
// latency-bound: ~14 dependent FP64 operations plus a table lookup and a MUFU).
template <typename T, int N> EB_HD void sigmoid_vec(const T (&g)[N], T (&out)[N]) {
#if defined(EB_SIGMOID_VEC)
  if constexpr (sizeof(T) == 8) {
    // the same arithmetic as sigmoid_t<double>, written stage by stage; table entries are fetched as soon as n is known
    const double* c = math_consts();
    const double magic = 6755399441055744.0;
    double t[N], r[N], p[N], tj[N];
    int n[N];
#pragma unroll
    for (int i = 0; i < N; ++i) t[i] = fma(-g[i], c[0], magic);
#pragma unroll
    for (int i = 0; i < N; ++i) { n[i] = dbl_lo(t[i]); tj[i] = exp2_tab(n[i] & (kExpTabN - 1)); }
#pragma unroll
    for (int i = 0; i < N; ++i) r[i] = fma(-g[i], c[0], -(t[i] - magic));
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = fma(r[i], c[3], c[2]);
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = fma(p[i], r[i], c[1]);
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = fma(r[i], p[i], 1.0);
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = fma(dbl_add_exponent(tj[i], n[i] >> 11), p[i], 1.0);
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = rcp_ge1(p[i]);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const int hi = dbl_hi(g[i]);
      const bool big = (int)((unsigned)(hi & 0x7fffffff) + 0xfffffu) > (0x40862000 + 0xfffff);
      out[i] = big ? dbl_from(hi < 0 ? 0 : 0x3ff00000, 0) : p[i];
    }
    return;
  }
#endif
#pragma unroll
  for (int i = 0; i < N; ++i) out[i] = sigmoid_t<T>(g[i]);
}

// sqrt(x) for positive, normal, finite x well inside the exponent range (Box-Muller radii: 1e-16 < x < 1500): reciprocal
// square root seed (MUFU.RSQ64H), two coupled Newton steps on (g ~ sqrt x, h ~ 1 / (2 sqrt x)) and a final residual
// correction; branch-free, 11 FP64 instructions, within 1 ulp.  (CUDA's sqrt() adds a range check and an out-of-line slow
// path, which splits the basic block between the ten Box-Muller pairs of an iteration.)
EB_HD double sqrt_pos(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#else
  const double y = 1.0 / sqrt(x);
#endif
  double g = x * y, h = 0.5 * y;
  double r = fma(-g, h, 0.5);
  g = fma(g, r, g); h = fma(h, r, h);
  r = fma(-g, h, 0.5);
  g = fma(g, r, g); h = fma(h, r, h);
  return fma(fma(-g, g, x), h, g);
}

// l = sqrt(x) and li = 1 / l together, for positive x with 1e-290 < x < 1e290 (Cholesky pivots): the coupled Newton iteration of
// sqrt_pos carries h ~ 1 / (2 sqrt x) anyway, so the reciprocal pivot costs one more FMA instead of an IEEE division whose
// dependent chain (seed, refinements, slow-path test) is as long as the square root's.  Both within 2 ulp.
EB_HD void sqrt_rsqrt_pos(double x, double* l, double* li) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#else
  const double y = 1.0 / sqrt(x);
#endif
  double g = x * y, h = 0.5 * y;
  double r = fma(-g, h, 0.5);
  g = fma(g, r, g); h = fma(h, r, h);
  r = fma(-g, h, 0.5);
  g = fma(g, r, g); h = fma(h, r, h);
  r = fma(-g, h, 0.5);                       // residual of the pair: both take their last correction from it
  *l = fma(fma(-g, g, x), h, g);
  *li = 2.0 * fma(h, r, h);
}

// sin(2 pi u), cos(2 pi u) for u in [0, 1): q = rint(4 u), r = 2 u - q / 2 in [-1/4, 1/4] (exact), Taylor polynomials of
// sin(pi r) / cos(pi r) in r^2 (truncation 5e-17 / 2e-18), quadrant fix-up by integer sign flips.  Branch-free.
EB_HD void sincos2pi_f64(double u, double* sn, double* cs) {
  const double* c = math_consts();
  const double magic = 6755399441055744.0;
  const double t = fma(u, 4.0, magic);
  const int iq = dbl_lo(t);
  const double r = fma(t - magic, -0.5, u + u);
  const double z = r * r;
  double ps = fma(z, c[13], c[12]);
  double pc = fma(z, c[21], c[20]);
  ps = fma(ps, z, c[11]); pc = fma(pc, z, c[19]);
  ps = fma(ps, z, c[10]); pc = fma(pc, z, c[18]);
  ps = fma(ps, z, c[9]);  pc = fma(pc, z, c[17]);
  ps = fma(ps, z, c[8]);  pc = fma(pc, z, c[16]);
  ps = fma(ps, z, c[7]);  pc = fma(pc, z, c[15]);
  ps = fma(ps, z, c[6]);  pc = fma(pc, z, c[14]);
  const double S = ps * r, C = fma(pc, z, 1.0);
  // quadrant iq mod 4: 0 (S, C), 1 (C, -S), 2 (-S, -C), 3 (-C, S)
  const bool swap = (iq & 1) != 0;
  const double a = swap ? C : S, b = swap ? S : C;
  *sn = dbl_from(dbl_hi(a) ^ ((iq & 2) << 30), dbl_lo(a));
  *cs = dbl_from(dbl_hi(b) ^ (((iq + 1) & 2) << 30), dbl_lo(b));
}

// cos/sin(2 pi u)
template <typename T> EB_HD void sincos2pi(T u, T* s, T* c);
template <> EB_HD void sincos2pi<float>(float u, float* s, float* c) {
#if defined(__CUDA_ARCH__)
  __sincosf(6.283185307179586f * u, s, c);   // MUFU.SIN / MUFU.COS: the argument lies in (0, 2 pi), absolute error ~5e-7
#else
  *s = sinf(6.283185307179586f * u); *c = cosf(6.283185307179586f * u);
#endif
}
template <> EB_HD void sincos2pi<double>(double u, double* s, double* c) { sincos2pi_f64(u, s, c); }

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
