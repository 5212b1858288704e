// Shared device/host helpers for the chain-batched kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <limits>

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

// sigmoid exactly as the reference evaluates it: 1 / (1 + exp(-g))  (torch.sigmoid, eeyore/models/mlp.py:48-49)
template <typename T> EB_HD T sigmoid_t(T g) { return T(1) / (T(1) + exp_t<T>(-g)); }

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
template <int G, typename T> EB_D T group_allreduce(T v) {
#pragma unroll
  for (int m = G / 2; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}
#endif

}  // namespace eb
