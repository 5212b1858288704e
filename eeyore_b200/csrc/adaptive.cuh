// Adaptive random-walk Metropolis samplers, one WARP per chain, all iterations of a run in one launch:
//   AM  -- adaptive Metropolis (Haario et al.), eeyore/samplers/am.py:62-107
//   RAM -- robust adaptive Metropolis (Vihola),  eeyore/samplers/ram.py:39-70
// Both propose theta' = theta + (factor) z and accept with log u < lt' - lt; they differ in how the P x P proposal factor
// adapts.  The per-chain matrices live in shared memory (lower triangles, leading dimension P + 1), lane i owns element i
// of every vector and row i of every matrix update; the Cholesky factorisation is the in-warp routine of smmala.cuh.
// The 32 lanes split the data rows of the log-target evaluation (eval_target<..., G = 32>).
//
// AM state per chain (in/out, so that a run can be continued): running mean [P], sum of theta theta^T [P, P], covariance
// estimate [P, P], number of accepted moves.  RAM state: the lower Cholesky factor [P, P].
// A failed factorisation (torch.linalg.cholesky raises in the reference) freezes the chain and records 1 + iteration in
// status[chain]; the Python layer turns that into the reference's RuntimeError.
#pragma once
#include "smmala.cuh"

namespace eb {

constexpr int kAdWarps = 4;
enum AdaptKind { ADAPT_AM = 0, ADAPT_RAM = 1 };

struct AdaptArgs {
  double p0, p1, p2;   // AM: l, b, c;  RAM: a (target acceptance), g (decay exponent), unused
  int t0;              // AM: t0
  long iter0;          // counter.idx at the first iteration of this launch
  void* state;         // [C, state_len] of T
  const void* cov0;    // AM: [P, P] initial covariance (shared by the chains)
  int32_t* status;     // [C]
};

template <class NET> constexpr int adapt_state_len(int kind) {
  return kind == ADAPT_AM ? NET::P + 2 * NET::P * NET::P + 1 : NET::P * NET::P;
}

template <typename T, class NET> struct AdWarpMem {
  static constexpr int LD = NET::P + 1;
  T L[NET::P * LD];     // factor used by the proposal
  T C[NET::P * LD];     // AM: covariance estimate;  RAM: the matrix to factorise
  T S[NET::P * LD];     // AM: sum of theta theta^T
  T dinv[32];
  T vec[32];
};

template <typename T> EB_D T philox_uniform2(RngKey key, uint32_t chain, uint32_t iter) {
  U4 w = philox4x32_10(U4{0u, iter, chain, 2u}, key.k0, key.k1);
  if constexpr (sizeof(T) == 8) return Uni<double>::from(w.x, w.y);
  else return Uni<float>::from(w.x);
}

// y_i = sum_{j <= i} L_ij v_j ; lane i holds v_i on entry and y_i on return
template <typename T, int P, int LD> EB_D T warp_lower_matvec(const T* L, T v, T* scratch) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  if (lane < P) scratch[lane] = v;
  __syncwarp();
  T y = T(0);
  if (lane < P)
    for (int j = 0; j <= lane; ++j) y = fma_t<T>(L[lane * LD + j], scratch[j], y);
  __syncwarp();
  return y;
}

template <typename T, class NET, int KIND>
__global__ void __launch_bounds__(kAdWarps * 32) adaptive_kernel(const ChainArgs<T> a, const AdaptArgs ad) {
  constexpr int P = NET::P, LD = P + 1;
  static_assert(P <= 32, "one lane per vector element");
  extern __shared__ __align__(16) unsigned char smem[];
  const SmemLayout<T, NET> lay(a.n_rows, kAdWarps, false);
  const DataView<T> d = stage_data<T, NET>(smem, lay, a);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  AdWarpMem<T, NET>& sm = reinterpret_cast<AdWarpMem<T, NET>*>(smem + align16(lay.total))[warp];
  long chain = (long)blockIdx.x * kAdWarps + warp;
  const bool live = chain < a.n_chains;
  if (!live) chain = a.n_chains - 1;
  const uint32_t gchain = a.chain0 + (uint32_t)chain;
  constexpr int SLEN = adapt_state_len<NET>(KIND);
  T* st = reinterpret_cast<T*>(ad.state) + (size_t)chain * SLEN;
  const T* cov0 = reinterpret_cast<const T*>(ad.cov0);

  // ---- state in: lane i owns theta_i, mean_i and row i of the matrices ------------------------------------------------
  T th_l = lane < P ? a.theta[chain * a.st_c + lane * a.st_p] : T(0);
  T lt_c = a.target[chain];
  T mean_l = T(0), n_accepted = T(0);
  if (lane < P) {
    if constexpr (KIND == ADAPT_AM) {
      mean_l = st[lane];
      for (int j = 0; j <= lane; ++j) {
        sm.S[lane * LD + j] = st[P + lane * P + j];
        sm.C[lane * LD + j] = st[P + P * P + lane * P + j];
      }
    } else {
      for (int j = 0; j <= lane; ++j) sm.L[lane * LD + j] = st[lane * P + j];
    }
  }
  if constexpr (KIND == ADAPT_AM) n_accepted = st[P + 2 * P * P];
  __syncwarp();
  bool dead = ad.status[chain] != 0;
  uint32_t n_acc = 0;

  for (long t = 0; t < a.n_iters; ++t) {
    const long idx = ad.iter0 + t;
    // ---- noise ---------------------------------------------------------------------------------------------------------
    T zl = T(0), u1 = T(0.5), u2;
    if (a.rng_mode == 0) {
      constexpr int PER = sizeof(T) == 8 ? 2 : 4;
      if (lane < P) {
        U4 w = philox4x32_10(U4{(uint32_t)(lane / PER), a.iter0 + (uint32_t)t, gchain, 0u}, a.key.k0, a.key.k1);
        T v[4];
        if constexpr (sizeof(T) == 8) {
          box_muller<T>(Uni<double>::from(w.x, w.y), Uni<double>::from(w.z, w.w), &v[0], &v[1]);
          v[2] = v[3] = T(0);
        } else {
          box_muller<T>(Uni<float>::from(w.x), Uni<float>::from(w.y), &v[0], &v[1]);
          box_muller<T>(Uni<float>::from(w.z), Uni<float>::from(w.w), &v[2], &v[3]);
        }
        const int k = lane % PER;
        zl = k == 0 ? v[0] : (k == 1 ? v[1] : (k == 2 ? v[2] : v[3]));
      }
      u2 = philox_uniform<T>(a.key, gchain, a.iter0 + (uint32_t)t);
      if constexpr (KIND == ADAPT_AM) u1 = philox_uniform2<T>(a.key, gchain, a.iter0 + (uint32_t)t);
    } else {
      if (lane < P) zl = a.z_tape[((size_t)t * a.n_chains + chain) * P + lane];
      if constexpr (KIND == ADAPT_AM) {
        u1 = a.u_tape[((size_t)t * 2 + 0) * a.n_chains + chain];
        u2 = a.u_tape[((size_t)t * 2 + 1) * a.n_chains + chain];
      } else {
        u2 = a.u_tape[(size_t)t * a.n_chains + chain];
      }
    }
    // ---- proposal ------------------------------------------------------------------------------------------------------
    T step_l;
    if constexpr (KIND == ADAPT_AM) {
      const bool adaptive = (idx + 1 > ad.t0) && !(u1 < (T)ad.p0);                 // am.py:69-76
      if (adaptive && !dead) {
        if (lane < P)
          for (int j = 0; j <= lane; ++j) sm.L[lane * LD + j] = sm.C[lane * LD + j];
        __syncwarp();
        T logdet;
        if (!warp_chol_inplace<T, P, LD>(sm.L, sm.dinv, logdet)) {                 // torch.linalg.cholesky(self.cov) raises
          dead = true;
          if (lane == 0 && live) ad.status[chain] = (int32_t)(1 + idx);
        }
        step_l = (T)ad.p1 * warp_lower_matvec<T, P, LD>(sm.L, zl, sm.vec);
      } else {
        step_l = (T)ad.p2 * zl;
      }
    } else {
      step_l = warp_lower_matvec<T, P, LD>(sm.L, zl, sm.vec);                      // ram.py:46
    }
    const T prop_l = th_l + step_l;
    T lt_p;
    {
      T th_p[P], g_unused[1];
#pragma unroll
      for (int j = 0; j < P; ++j) th_p[j] = __shfl_sync(0xffffffffu, prop_l, j);
      eval_target<T, NET, 32, false>(d, lane, th_p, lt_p, g_unused);
    }
    const T log_rate = lt_p - lt_c;
    const bool acc = !dead && (log_t<T>(u2) < log_rate);
    if (acc) {
      ++n_acc;
      th_l = prop_l;
      lt_c = lt_p;
      if constexpr (KIND == ADAPT_AM)
        if (idx > 0) n_accepted += T(1);                                            // am.py:87-88
    }
    // ---- adaptation ----------------------------------------------------------------------------------------------------
    if (!dead) {
      if constexpr (KIND == ADAPT_AM) {
        const T k = (T)(idx + 1);
        mean_l = ((k - T(1)) * mean_l + th_l) / k;                                  // stats/recursive_mean.py
        __syncwarp();
        if (lane < P) sm.vec[lane] = th_l;
        __syncwarp();
        if (lane < P)
          for (int j = 0; j <= lane; ++j) sm.S[lane * LD + j] += th_l * sm.vec[j];  // cov_sum + ger(theta, theta)
        __syncwarp();
        if (idx + 1 >= ad.t0) {                                                     // am.py:96-102
          if (lane < P) sm.vec[lane] = mean_l;
          __syncwarp();
          if (lane < P) {
            const T kk = (T)idx;
            for (int j = 0; j <= lane; ++j)
              sm.C[lane * LD + j] = n_accepted == T(0) ? cov0[lane * P + j]
                                                       : (sm.S[lane * LD + j] - (kk + T(1)) * (mean_l * sm.vec[j])) / kk;
          }
          __syncwarp();
        }
      } else {
        // ram.py:61-66: chol( L (I + h (min(1, exp(log_rate)) - a) z z^T / |z|^2) L^T ) = chol(L L^T + coef w w^T), w = L z
        const double h = fmin(1.0, (double)P * pow((double)(idx + 1), -ad.p1));
        const T ex = exp_t<T>(log_rate);
        const T alpha = ex < T(1) ? ex : T(1);                                      // python min(1, nan) == 1
        const T zz = warp_sum<T>(zl * zl);
        const T coef = (T)h * (alpha - (T)ad.p0) / zz;
        __syncwarp();
        if (lane < P) sm.vec[lane] = step_l;
        __syncwarp();
        if (lane < P)
          for (int j = 0; j <= lane; ++j) {
            T m = coef * (step_l * sm.vec[j]);
            for (int k = 0; k <= j; ++k) m = fma_t<T>(sm.L[lane * LD + k], sm.L[j * LD + k], m);
            sm.C[lane * LD + j] = m;
          }
        __syncwarp();
        if (lane < P)
          for (int j = 0; j <= lane; ++j) sm.L[lane * LD + j] = sm.C[lane * LD + j];
        __syncwarp();
        T logdet;
        if (!warp_chol_inplace<T, P, LD>(sm.L, sm.dinv, logdet)) {
          dead = true;
          if (lane == 0 && live) ad.status[chain] = (int32_t)(1 + idx);
        }
      }
    }
    // ---- saved state -----------------------------------------------------------------------------------------------------
    if (t >= a.n_burnin && (t - a.n_burnin) % a.thin == 0 && live) {
      const long s = (t - a.n_burnin) / a.thin;
      if (lane < P && a.out_samples) a.out_samples[s * a.ss_i + chain * a.ss_c + lane * a.ss_p] = th_l;
      if (lane == 0) {
        if (a.out_target) a.out_target[s * a.n_chains + chain] = lt_c;
        if (a.out_acc) a.out_acc[s * a.n_chains + chain] = acc ? 1 : 0;
      }
    }
  }
  // ---- state out ----------------------------------------------------------------------------------------------------------
  if (live) {
    if (lane < P) {
      a.theta[chain * a.st_c + lane * a.st_p] = th_l;
      if constexpr (KIND == ADAPT_AM) {
        st[lane] = mean_l;
        for (int j = 0; j <= lane; ++j) {
          st[P + lane * P + j] = sm.S[lane * LD + j];
          st[P + P * P + lane * P + j] = sm.C[lane * LD + j];
        }
      } else {
        for (int j = 0; j <= lane; ++j) st[lane * P + j] = sm.L[lane * LD + j];
      }
    }
    if (lane == 0) {
      a.target[chain] = lt_c;
      if (a.acc_count) a.acc_count[chain] += n_acc;
      if constexpr (KIND == ADAPT_AM) st[P + 2 * P * P] = n_accepted;
    }
  }
}

template <typename T, class NET> cudaError_t launch_adaptive(int kind, const ChainArgs<T>& a, const AdaptArgs& ad, cudaStream_t st) {
  const SmemLayout<T, NET> lay(a.n_rows, kAdWarps, false);
  const size_t smem = align16(lay.total) + kAdWarps * sizeof(AdWarpMem<T, NET>);
  const long blocks = (a.n_chains + kAdWarps - 1) / kAdWarps;
  cudaError_t e;
  if (kind == ADAPT_AM) {
    auto kern = adaptive_kernel<T, NET, ADAPT_AM>;
    e = reserve_smem(kern, smem);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)blocks, kAdWarps * 32, smem, st>>>(a, ad);
  } else {
    auto kern = adaptive_kernel<T, NET, ADAPT_RAM>;
    e = reserve_smem(kern, smem);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)blocks, kAdWarps * 32, smem, st>>>(a, ad);
  }
  return cudaGetLastError();
}

}  // namespace eb
