// On-device chain diagnostics: mean, sample covariance, initial-sequence Monte Carlo covariance (INSE), multivariate
// ESS and autocorrelation, batched over chains (one CTA per chain, chain staged in shared memory).
//
// Replaces (reference paths):
//   eeyore/stats/cov.py:5-15           sample covariance (n-1 denominator)
//   eeyore/linalg/is_pos_def.py:3-11   "exactly symmetric and Cholesky succeeds"
//   eeyore/stats/inse_mc_cov.py:9-83   INSE estimator (adjust=False): the python double loop of torch.ger outer products
//   eeyore/stats/multi_ess.py:6-14     n (det cov / det inse)^(1/p)
// ACF is builder-defined (kanga is absent from the reference snapshot).
#include <cuda_runtime.h>
#include <string>
#include "common.cuh"
#include "../../include/eeyore_b200.h"

namespace eb {

constexpr int kStatThreads = 256;
constexpr int kTile = 4;  // each thread owns a 4x4 tile of a lagged cross-product matrix

template <typename T> struct StatsArgs {
  const T* x;
  long s_iter, s_chain, s_param;
  int n, P;
  long C;
  T* out_mean;   // [C,P]
  T* out_cov;    // [C,P,P]
  T* out_inse;   // [C,P,P]
  T* out_ess;    // [C]
  int* out_status;  // [C] 0 ok, 1 = not enough samples (inse_mc_cov.py:44-45), 2 = non-finite input
  int* out_lags;    // [C,2] (sn, last accepted m)
  int max_lag;
  T* out_acf;    // [C, max_lag+1, P]
  T* scratch;    // global fallback for the centred chain ([grid, n, PS]) when it does not fit shared memory
  int use_scratch;
};

__host__ __device__ inline int stat_ps(int P) { return ((P + kTile - 1) / kTile) * kTile + 1; }  // odd row stride

// Lagged cross-product A[a][b] = sum_{i < n-lag} xc[i][a] * xc[i+lag][b], all threads of the CTA cooperate:
// thread -> (tile, slice of the i range); per-slice partial tiles are reduced through shared memory.
template <typename T>
__device__ void lag_product(const T* __restrict__ xc, int n, int P, int PS, int lag, T* part, T* out, int tiles_1d,
                            int slices) {
  const int ntiles = tiles_1d * tiles_1d;
  const int tid = threadIdx.x;
  const int PP = tiles_1d * kTile;
  if (tid < ntiles * slices) {
    const int tile = tid % ntiles, slice = tid / ntiles;
    const int a0 = (tile / tiles_1d) * kTile, b0 = (tile % tiles_1d) * kTile;
    const int len = n - lag;
    const int chunk = (len + slices - 1) / slices;
    const int i0 = slice * chunk, i1 = min(len, i0 + chunk);
    T acc[kTile][kTile];
#pragma unroll
    for (int r = 0; r < kTile; ++r)
#pragma unroll
      for (int c = 0; c < kTile; ++c) acc[r][c] = T(0);
    for (int i = i0; i < i1; ++i) {
      T va[kTile], vb[kTile];
#pragma unroll
      for (int r = 0; r < kTile; ++r) { va[r] = xc[(size_t)i * PS + a0 + r]; vb[r] = xc[(size_t)(i + lag) * PS + b0 + r]; }
#pragma unroll
      for (int r = 0; r < kTile; ++r)
#pragma unroll
        for (int c = 0; c < kTile; ++c) acc[r][c] = fma_t<T>(va[r], vb[c], acc[r][c]);
    }
#pragma unroll
    for (int r = 0; r < kTile; ++r)
#pragma unroll
      for (int c = 0; c < kTile; ++c) part[(size_t)slice * PP * PP + (a0 + r) * PP + b0 + c] = acc[r][c];
  }
  __syncthreads();
  for (int e = tid; e < P * P; e += blockDim.x) {
    const int a = e / P, b = e % P;
    T s = T(0);
    for (int sl = 0; sl < slices; ++sl) s += part[(size_t)sl * PP * PP + a * PP + b];
    out[e] = s;
  }
  __syncthreads();
}

// Warp-level Cholesky of the P x P matrix m (row-major, stride P) into w; returns false when a pivot is not
// positive (or NaN), i.e. torch.linalg.cholesky would raise (is_pos_def.py:5-9).  log_det (optional) = 2 sum log L_jj.
template <typename T> __device__ bool warp_cholesky(const T* m, T* w, int P, T* det_out) {
  const int lane = threadIdx.x & 31;
  for (int e = lane; e < P * P; e += 32) w[e] = m[e];
  __syncwarp();
  bool ok = true;
  T det = T(1);
  for (int j = 0; j < P; ++j) {
    T d = w[j * P + j];
    for (int k = 0; k < j; ++k) d -= w[j * P + k] * w[j * P + k];
    if (!(d > T(0))) { ok = false; break; }
    const T l = sqrt_t<T>(d);
    det *= d;
    __syncwarp();
    for (int i = j + 1 + lane; i < P; i += 32) {
      T s = w[i * P + j];
      for (int k = 0; k < j; ++k) s -= w[i * P + k] * w[j * P + k];
      w[i * P + j] = s / l;
    }
    if (lane == 0) w[j * P + j] = l;
    __syncwarp();
  }
  if (det_out) *det_out = det;
  return ok;
}

// Warp-level determinant by LU with partial pivoting (torch.det, inse_mc_cov.py:47,66 and multi_ess.py:9-12).
template <typename T> __device__ T warp_det_lu(const T* m, T* w, int P) {
  const int lane = threadIdx.x & 31;
  for (int e = lane; e < P * P; e += 32) w[e] = m[e];
  __syncwarp();
  T det = T(1);
  for (int k = 0; k < P; ++k) {
    // pivot search (all lanes redundantly; P <= 32)
    int piv = k;
    T best = fabs(w[k * P + k]);
    for (int i = k + 1; i < P; ++i) {
      const T v = fabs(w[i * P + k]);
      if (v > best) { best = v; piv = i; }
    }
    __syncwarp();
    if (piv != k) {
      for (int c = lane; c < P; c += 32) { const T t = w[k * P + c]; w[k * P + c] = w[piv * P + c]; w[piv * P + c] = t; }
      det = -det;
    }
    __syncwarp();
    const T pv = w[k * P + k];
    det *= pv;
    if (pv == T(0) || pv != pv) break;
    for (int i = k + 1 + lane; i < P; i += 32) {
      const T f = w[i * P + k] / pv;
      for (int c = k + 1; c < P; ++c) w[i * P + c] -= f * w[k * P + c];
    }
    __syncwarp();
  }
  return det;
}

template <typename T> __global__ void __launch_bounds__(kStatThreads) chain_stats_kernel(const StatsArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = a.n, P = a.P, PS = stat_ps(P);
  const int tiles_1d = (P + kTile - 1) / kTile, ntiles = tiles_1d * tiles_1d, PP = tiles_1d * kTile;
  const int slices = max(1, min(kStatThreads / ntiles, 16));
  const int tid = threadIdx.x;
  // carve: matrices first, then the (optional) chain buffer
  T* part = reinterpret_cast<T*>(smem_raw);            // [slices, PP, PP]
  T* A0 = part + (size_t)slices * PP * PP;             // lag product (even lag)
  T* A1 = A0 + P * P;                                  // lag product (odd lag)
  T* Sig = A1 + P * P;                                 // running INSE estimate
  T* Sig1 = Sig + P * P;                               // candidate
  T* Cov = Sig1 + P * P;
  T* work = Cov + P * P;                               // factorisation workspace
  T* mean = work + P * P;                              // [PP]
  T* ctrl = mean + PP;                                 // [4] broadcast slots
  T* xs_sh = ctrl + 4;
  const T inv_n = T(1) / T(n);

  for (long c = blockIdx.x; c < a.C; c += gridDim.x) {
    T* xc = a.use_scratch ? a.scratch + (size_t)blockIdx.x * n * PS : xs_sh;
    // ---- stage the chain (zero-padded columns), mean, centring ------------------------------------------------
    for (int e = tid; e < n * PS; e += blockDim.x) {
      const int i = e / PS, j = e % PS;
      xc[e] = (j < P) ? a.x[i * a.s_iter + c * a.s_chain + j * a.s_param] : T(0);
    }
    __syncthreads();
    {
      // column sums: thread -> (column, slice)
      const int cs = max(1, kStatThreads / PP);
      if (tid < PP * cs) {
        const int j = tid % PP, sl = tid / PP;
        const int chunk = (n + cs - 1) / cs;
        T s = T(0);
        for (int i = sl * chunk; i < min(n, (sl + 1) * chunk); ++i) s += xc[(size_t)i * PS + j];
        part[sl * PP + j] = s;
      }
      __syncthreads();
      if (tid < PP) {
        T s = T(0);
        for (int sl = 0; sl < cs; ++sl) s += part[sl * PP + tid];
        mean[tid] = s * inv_n;
      }
      __syncthreads();
      for (int e = tid; e < n * PS; e += blockDim.x) {
        const int j = e % PS;
        if (j < P) xc[e] -= mean[j];
      }
      __syncthreads();
    }
    if (a.out_mean) for (int j = tid; j < P; j += blockDim.x) a.out_mean[c * P + j] = mean[j];

    // ---- autocorrelation (builder-defined) -----------------------------------------------------
    if (a.out_acf) {
      const int K = a.max_lag + 1;
      for (int e = tid; e < K * P; e += blockDim.x) {
        const int k = e / P, j = e % P;
        T num = T(0), den = T(0);
        for (int t = 0; t < n; ++t) {
          const T v = xc[(size_t)t * PS + j];
          den = fma_t<T>(v, v, den);
          if (t + k < n) num = fma_t<T>(v, xc[(size_t)(t + k) * PS + j], num);
        }
        a.out_acf[(c * K + k) * P + j] = num / den;
      }
    }
    if (!a.out_cov && !a.out_inse && !a.out_ess) { __syncthreads(); continue; }

    // ---- lag 0: covariance and gamma_0 -------------------------------------------------------------------------
    lag_product<T>(xc, n, P, PS, 0, part, A0, tiles_1d, slices);
    for (int e = tid; e < P * P; e += blockDim.x) {
      Cov[e] = A0[e] / T(n - 1);                       // cov.py:13-15
      if (a.out_cov) a.out_cov[c * P * P + e] = Cov[e];
    }
    __syncthreads();
    if (!a.out_inse && !a.out_ess) continue;

    // ---- INSE (inse_mc_cov.py:20-73) ---------------------------------------------------------------------------
    const int ub = n / 2;
    int sn = ub, m_last = -1, status = 0;
    T last_det = T(0);
    bool phase2 = false;
    for (int m = 0; m < ub; ++m) {
      if (m > 0) lag_product<T>(xc, n, P, PS, 2 * m, part, A0, tiles_1d, slices);
      if (2 * m + 1 < n) lag_product<T>(xc, n, P, PS, 2 * m + 1, part, A1, tiles_1d, slices);
      T* dst = phase2 ? Sig1 : Sig;
      for (int e = tid; e < P * P; e += blockDim.x) {
        const int r = e / P, q = e % P, et = q * P + r;
        // Gam = sym(gam0 + gam1), :32-33.  Written with pure additions before the scaling so that FMA contraction
        // cannot round element (r, q) and (q, r) differently: the reference's is_pos_def demands EXACT symmetry.
        const T se = A0[e] + A1[e], st = A0[et] + A1[et];
        const T gam = (se + st) * (T(0.5) * inv_n);
        const T g0 = A0[e] * inv_n;
        if (m == 0) dst[e] = T(2) * gam - g0;                          // :35-36 (A0 is exactly symmetric at lag 0)
        else dst[e] = Sig[e] + T(2) * gam;                             // :37-38, :62
      }
      __syncthreads();
      if (tid < 32) {
        if (!phase2) {
          T det;
          const bool pd = warp_cholesky<T>(Sig, work, P, &det);         // is_pos_def(Sig), :40
          if (tid == 0) ctrl[0] = pd ? T(1) : T(0);
          if (pd) {
            const T dlu = warp_det_lu<T>(Sig, work, P);                  // last_dtm = det(Sig), :47
            if (tid == 0) ctrl[1] = dlu;
          }
        } else {
          const T dlu = warp_det_lu<T>(Sig1, work, P);                   // :66
          if (tid == 0) ctrl[1] = dlu;
        }
      }
      __syncthreads();
      if (!phase2) {
        if (ctrl[0] != T(0)) { sn = m; m_last = m; last_det = ctrl[1]; phase2 = true; }
      } else {
        const T cur = ctrl[1];
        if (!(cur > last_det)) break;                                  // current_dtm <= last_dtm -> break, :68-69
        for (int e = tid; e < P * P; e += blockDim.x) Sig[e] = Sig1[e];
        last_det = cur;
        m_last = m;
      }
      __syncthreads();
    }
    if (sn > ub - 1) status = 1;                                        // 'Not enough samples', :44-45
    if (a.out_inse) for (int e = tid; e < P * P; e += blockDim.x) a.out_inse[c * P * P + e] = Sig[e];
    // ---- multi-ESS (multi_ess.py:9-14) ---------------------------------------------------------------------------
    if (tid < 32) {
      const T dcov = warp_det_lu<T>(Cov, work, P);
      if (tid == 0) {
        const double ratio = (double)dcov / (double)last_det;
        const T ess = (T)((double)n * pow(ratio, 1.0 / (double)P));
        if (a.out_ess) a.out_ess[c] = status == 0 ? ess : qnan<T>();
        if (a.out_status) a.out_status[c] = status;
        if (a.out_lags) { a.out_lags[2 * c] = sn; a.out_lags[2 * c + 1] = m_last; }
      }
    }
    __syncthreads();
  }
}

template <typename T> size_t stats_smem_fixed(int P) {
  const int tiles_1d = (P + kTile - 1) / kTile, ntiles = tiles_1d * tiles_1d, PP = tiles_1d * kTile;
  int slices = kStatThreads / ntiles;
  if (slices < 1) slices = 1;
  if (slices > 16) slices = 16;
  size_t part = (size_t)slices * PP * PP;
  const size_t cs = (size_t)(kStatThreads / PP > 0 ? kStatThreads / PP : 1) * PP;
  if (cs > part) part = cs;
  return sizeof(T) * (part + 6 * (size_t)P * P + PP + 4);
}

template <typename T> cudaError_t launch_stats(StatsArgs<T> a, cudaStream_t st) {
  int dev = 0, sms = 0, max_smem = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const int PS = stat_ps(a.P);
  const size_t fixed = stats_smem_fixed<T>(a.P);
  const size_t chain_bytes = sizeof(T) * (size_t)a.n * PS;
  size_t smem = fixed + chain_bytes;
  long grid = a.C < (long)sms * 8 ? a.C : (long)sms * 8;
  a.use_scratch = 0;
  a.scratch = nullptr;
  cudaError_t e;
  if (smem > (size_t)max_smem) {  // chain does not fit on chip: keep the centred chain in global memory (L2)
    smem = fixed;
    grid = a.C < (long)sms * 2 ? a.C : (long)sms * 2;
    a.use_scratch = 1;
    e = cudaMallocAsync((void**)&a.scratch, chain_bytes * grid, st);
    if (e != cudaSuccess) return e;
  }
  e = cudaFuncSetAttribute(chain_stats_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  chain_stats_kernel<T><<<(unsigned)grid, kStatThreads, smem, st>>>(a);
  e = cudaGetLastError();
  if (a.use_scratch) cudaFreeAsync(a.scratch, st);
  return e;
}

}  // namespace eb

using namespace eb;

extern "C" {

// defined in capi.cu
int eeyore_b200_set_error_(int code, const char* msg);

int eeyore_b200_chain_stats(int dtype, int64_t n_chains, int64_t n_samples, int n_params, const void* samples,
                            int64_t ss_iter, int64_t ss_chain, int64_t ss_param, void* out_mean, void* out_cov,
                            void* out_inse, void* out_ess, int32_t* out_status, int32_t* out_lags, int max_lag,
                            void* out_acf, void* stream) {
  if (!samples || n_chains < 1 || n_samples < 2 || n_params < 1)
    return eeyore_b200_set_error_(EEYORE_B200_EINVAL, "chain_stats: bad sizes or null samples");
  if (n_params > 32) return eeyore_b200_set_error_(EEYORE_B200_EUNSUPPORTED, "chain_stats: at most 32 parameters per chain");
  if (out_acf && (max_lag < 0 || max_lag >= n_samples))
    return eeyore_b200_set_error_(EEYORE_B200_EINVAL, "acf: max_lag must be in [0, n_samples)");
  cudaError_t e;
  if (dtype == EEYORE_B200_F64) {
    StatsArgs<double> a{(const double*)samples, ss_iter, ss_chain, ss_param, (int)n_samples, n_params, n_chains,
                        (double*)out_mean, (double*)out_cov, (double*)out_inse, (double*)out_ess, out_status, out_lags,
                        max_lag, (double*)out_acf, nullptr, 0};
    e = launch_stats<double>(a, (cudaStream_t)stream);
  } else if (dtype == EEYORE_B200_F32) {
    StatsArgs<float> a{(const float*)samples, ss_iter, ss_chain, ss_param, (int)n_samples, n_params, n_chains,
                       (float*)out_mean, (float*)out_cov, (float*)out_inse, (float*)out_ess, out_status, out_lags,
                       max_lag, (float*)out_acf, nullptr, 0};
    e = launch_stats<float>(a, (cudaStream_t)stream);
  } else {
    return eeyore_b200_set_error_(EEYORE_B200_EINVAL, "dtype must be f32 or f64");
  }
  if (e != cudaSuccess) return eeyore_b200_set_error_(EEYORE_B200_ECUDA, cudaGetErrorString(e));
  return EEYORE_B200_OK;
}

}  // extern "C"
