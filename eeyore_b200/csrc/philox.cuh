// Philox4x32-10 (Salmon et al., SC'11) and the per-chain stream layout of the device samplers.
// Replaces the reference's global torch RNG draws: torch.randn(P) (eeyore/samplers/hmc.py:134),
// Normal.sample (eeyore/kernels/normalized_kernel.py:17-19) and torch.rand(1) (mala.py:66, hmc.py:148,
// metropolis_hastings.py:56).  The layout is pinned on the CPU side by oracle/philox.py.
//
//   key     = (seed lo, seed hi)
//   counter = (block j, iteration t, chain id c, kind)      kind 0 = normals, 1 = accept uniform
#pragma once
#include "common.cuh"

namespace eb {

struct U4 { uint32_t x, y, z, w; };

EB_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

EB_HD U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = mulhi32(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = mulhi32(M1, c.z), lo1 = M1 * c.z;
    U4 n;
    n.x = hi1 ^ c.y ^ k0; n.y = lo1; n.z = hi0 ^ c.w ^ k1; n.w = lo0;
    c = n;
    k0 += W0; k1 += W1;
  }
  return c;
}

template <typename T> struct Uni;
template <> struct Uni<double> {
  // 53-bit uniform in (0,1) from two words
  static EB_HD double from(uint32_t hi, uint32_t lo) {
    uint64_t v = ((uint64_t)hi << 32) | (uint64_t)lo;
    return ((double)(v >> 11) + 0.5) * 1.1102230246251565e-16;  // 2^-53
  }
};
template <> struct Uni<float> {
  // 23-bit uniform in (0,1): (k + 1/2) 2^-23, k < 2^23 -- every value is exact in fp32 (a 24-bit k + 1/2 would round to
  // even: neighbours collide and the largest one becomes exactly 1)
  static EB_HD float from(uint32_t w) { return ((float)(w >> 9) + 0.5f) * 1.1920928955078125e-7f; }  // 2^-23
};

// log of a uniform in (0, 1): a positive normal number in both precisions
template <typename T> EB_HD T log_unit_t(T u) { return log_t<T>(u); }
template <> EB_HD float log_unit_t<float>(float u) { return log_fast(u); }   // lg2.approx: u is a normal number in (0, 1)
template <> EB_HD double log_unit_t<double>(double u) { return log_pos_normal(u); }

// radius of the pair: -2 log u1 lies in [0, 75) for the 53-bit (24-bit) uniforms of Uni<>
template <typename T> EB_HD T radius_t(T u1) { return sqrt_t<T>(T(-2) * log_unit_t<T>(u1)); }
template <> EB_HD double radius_t<double>(double u1) {
  // Uni<double>::from rounds its largest value (2^53 - 1/2) 2^-53 to exactly 1: log = 0, radius 0 (sqrt_pos itself needs x > 0)
  const double x = -2.0 * log_pos_normal(u1);
  const double r = sqrt_pos(x);
  return (((dbl_hi(x) & 0x7fffffff) | dbl_lo(x)) == 0) ? 0.0 : r;
}

template <typename T> EB_HD void box_muller(T u1, T u2, T* z0, T* z1) {
  T r = radius_t<T>(u1);
  T s, c;
  sincos2pi<T>(u2, &s, &c);
  *z0 = r * c; *z1 = r * s;
}

struct RngKey { uint32_t k0, k1; };

// Fills z[0..P) for (chain, iteration).  P is a compile-time or runtime bound; VEC provides operator[].
template <typename T, int P, class VEC> EB_HD void philox_normals(VEC& z, RngKey key, uint32_t chain, uint32_t iter) {
  if constexpr (sizeof(T) == 8) {
#pragma unroll
    for (int j = 0; j < (P + 1) / 2; ++j) {
      U4 w = philox4x32_10(U4{(uint32_t)j, iter, chain, 0u}, key.k0, key.k1);
      T a, b;
      box_muller<T>(Uni<double>::from(w.x, w.y), Uni<double>::from(w.z, w.w), &a, &b);
      z[2 * j] = a;
      if (2 * j + 1 < P) z[2 * j + 1] = b;
    }
  } else {
#pragma unroll
    for (int j = 0; j < (P + 3) / 4; ++j) {
      U4 w = philox4x32_10(U4{(uint32_t)j, iter, chain, 0u}, key.k0, key.k1);
      T a, b, c, d;
      box_muller<T>(Uni<float>::from(w.x), Uni<float>::from(w.y), &a, &b);
      box_muller<T>(Uni<float>::from(w.z), Uni<float>::from(w.w), &c, &d);
      z[4 * j] = a;
      if (4 * j + 1 < P) z[4 * j + 1] = b;
      if (4 * j + 2 < P) z[4 * j + 2] = c;
      if (4 * j + 3 < P) z[4 * j + 3] = d;
    }
  }
}

template <typename T> EB_HD T philox_uniform(RngKey key, uint32_t chain, uint32_t iter) {
  U4 w = philox4x32_10(U4{0u, iter, chain, 1u}, key.k0, key.k1);
  if constexpr (sizeof(T) == 8) return Uni<double>::from(w.x, w.y);
  else return Uni<float>::from(w.x);
}

}  // namespace eb
