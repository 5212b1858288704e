// Per-chain MH / MALA / HMC draws on top of mlp_static.cuh (host+device code; thread mapping lives in
// chain_kernels.cuh).  Each function performs ONE reference `draw` for the full-batch case (num_batches == 1, cached
// current target / gradient reused) and returns the accept decision; the caller commits the proposal.
//
// Replaces (reference paths):
//   eeyore/samplers/metropolis_hastings.py:41-73   MetropolisHastings.draw
//   eeyore/samplers/mala.py:35-82                  MALA.kernel_mean / draw
//   eeyore/samplers/hmc.py:91-170                  HMC.hamiltonian / leapfrog / draw
//   eeyore/kernels/normalized_kernel.py:14-19      Normal sample / summed log_prob
#pragma once
#include "mlp_static.cuh"

namespace eb {

// The chain's current state (sample and gradient) lives in global memory (the run's in/out state arrays), element j at
// [j * stride] (chain-minor layout => coalesced); the current target value stays in a register.
template <typename T> struct Cur {
  T* th;
  T* g;
  long stride;
};

constexpr double kLogSqrt2Pi = 0.9189385332046727;  // log(sqrt(2 pi)), torch.distributions.Normal.log_prob

// Gradient accumulators: in registers, or (fp64, larger P) one shared-memory column per thread so that the kernel fits
// 128 registers and four CTAs stay resident per SM.
template <typename T, int P> struct RegVec {
  T v[P];
  EB_HD T& operator[](int j) { return v[j]; }
  EB_HD const T& operator[](int j) const { return v[j]; }
};
template <typename T> struct StridedVec {
  T* base;
  int stride;
  EB_HD T& operator[](int j) const { return base[j * stride]; }
};

// metropolis_hastings.py:41-73.  z: standard normals; prop_scale: kernel scale (default 1, :25-28).
template <typename T, class NET, int G, class TV>
EB_HD bool mh_draw(const DataView<T>& d, int sub, T prop_scale, bool symmetric, const Cur<T>& cur, T lt_cur,
                   const T (&z)[NET::P], T u, TV& thp, T& ltp) {
#pragma unroll
  for (int j = 0; j < NET::P; ++j) thp[j] = fma_t<T>(prop_scale, z[j], cur.th[j * cur.stride]);  // kernel.sample()
  int dummy = 0;
  eval_target<T, NET, G, false>(d, sub, thp, ltp, dummy);
  T log_rate = ltp - lt_cur;                                                                     // :49
  if (!symmetric) {                                                                              // :51-54
    const T inv2var = T(1) / (T(2) * prop_scale * prop_scale);
    const T lnorm = log_t<T>(prop_scale) + T(kLogSqrt2Pi);
    T lq_f = T(0), lq_b = T(0);
#pragma unroll
    for (int j = 0; j < NET::P; ++j) {
      const T a = thp[j] - cur.th[j * cur.stride];
      const T b = cur.th[j * cur.stride] - thp[j];
      lq_f += -(a * a) * inv2var - lnorm;
      lq_b += -(b * b) * inv2var - lnorm;
    }
    log_rate = log_rate - lq_f;
    log_rate = log_rate + lq_b;
  }
  return log_t<T>(u) < log_rate;                                                                 // :56 (NaN -> reject)
}

// mala.py:46-82.  sd = dtype(sqrt(step)) (mala.py:40); half_step = 0.5 * step (mala.py:36).
template <typename T, class NET, int G, class GV, class TV>
EB_HD bool mala_draw(const DataView<T>& d, int sub, T half_step, T sd, const Cur<T>& cur, T lt_cur,
                     const T (&z)[NET::P], T u, TV& thp, GV& gp, T& ltp) {
  const T inv2var = T(1) / (T(2) * (sd * sd));
  const T lnorm = log_t<T>(sd) + T(kLogSqrt2Pi);
  T lq_f = T(0);
#pragma unroll
  for (int j = 0; j < NET::P; ++j) {
    const T mean = fma_t<T>(half_step, cur.g[j * cur.stride], cur.th[j * cur.stride]);  // kernel_mean(current)
    thp[j] = fma_t<T>(sd, z[j], mean);                                                    // :53
    const T dd = thp[j] - mean;
    lq_f += -(dd * dd) * inv2var - lnorm;                                                 // log q(theta' | theta), :60
  }
  eval_target<T, NET, G, true>(d, sub, thp, ltp, gp);                                     // :55-56
  T lq_b = T(0);
#pragma unroll
  for (int j = 0; j < NET::P; ++j) {
    const T mean_p = fma_t<T>(half_step, gp[j], thp[j]);                                  // :62
    const T dd = cur.th[j * cur.stride] - mean_p;
    lq_b += -(dd * dd) * inv2var - lnorm;                                                 // log q(theta | theta'), :64
  }
  T log_rate = ltp - lt_cur;                                                              // :58
  log_rate = log_rate - lq_f;
  log_rate = log_rate + lq_b;
  return log_t<T>(u) < log_rate;                                                          // :66
}

// hmc.py:126-170 with leapfrog :100-124.  z is the momentum draw p0.  The momentum vector `p` is a StridedVec over shared
// memory (one copy per chain, lane `sub` owns entries j % G == sub), so that only theta' and the gradient accumulators
// occupy registers during an evaluation -- or a RegVec where one lane owns the chain and the register budget allows.
// The first gradient of the trajectory is the cached current gradient (the reference recomputes it at the same point
// with the same data, hmc.py:104, so the value is identical); num_steps further evaluations follow.
template <int G> EB_HD void group_sync() {
#if defined(__CUDA_ARCH__)
  if (G > 1) __syncwarp(group_mask<G>());
#endif
}

template <typename T, class NET, int G, class PV, class GV, class TV>
EB_HD bool hmc_draw(const DataView<T>& d, int sub, T eps, T half_eps, int num_steps, const Cur<T>& cur, T lt_cur,
                    const T (&z)[NET::P], PV& p, T u, TV& thp, GV& gp, T& ltp, T* rate_out = nullptr) {
  T kin = T(0);
#pragma unroll
  for (int j = 0; j < NET::P; ++j) kin = fma_t<T>(z[j], z[j], kin);
  const T h_cur = -lt_cur + T(0.5) * kin;                                                 // :137
#pragma unroll
  for (int j = 0; j < NET::P; ++j) {
    thp[j] = cur.th[j * cur.stride];
    if (j % G == sub) p[j] = fma_t<T>(half_eps, cur.g[j * cur.stride], z[j]);        // :105
  }
  group_sync<G>();
  ltp = lt_cur;
  if constexpr (G == 1) {   // one lane owns the chain: the first position update reads the momentum it has just formed
#pragma unroll
    for (int j = 0; j < NET::P; ++j) thp[j] = fma_t<T>(eps, p[j], thp[j]);           // :110
  }
  for (int s = 0; s < num_steps; ++s) {
    if constexpr (G > 1) {
#pragma unroll
      for (int j = 0; j < NET::P; ++j) thp[j] = fma_t<T>(eps, p[j], thp[j]);         // :110, :117
      group_sync<G>();
    }
    // the target value matters at the end of the trajectory only (:141); the inner steps evaluate the gradient alone
    if (s == num_steps - 1) eval_target<T, NET, G, true, true>(d, sub, thp, ltp, gp);     // :118
    else eval_target<T, NET, G, true, false>(d, sub, thp, ltp, gp);                       // :113
    const T w = (s == num_steps - 1) ? half_eps : eps;                                    // :114, :119
    if constexpr (G == 1) {   // momentum and the next position update in one pass over the parameters (no shared-memory
                              // round trip of the momentum between the two)
      const bool more = s + 1 < num_steps;
#pragma unroll
      for (int j = 0; j < NET::P; ++j) {
        const T pj = fma_t<T>(w, gp[j], p[j]);
        p[j] = pj;
        if (more) thp[j] = fma_t<T>(eps, pj, thp[j]);                                 // :117
      }
    } else {
#pragma unroll
      for (int j = 0; j < NET::P; ++j)
        if (j % G == sub) p[j] = fma_t<T>(w, gp[j], p[j]);
      group_sync<G>();
    }
  }
  // momentum negation (:122) leaves the kinetic energy unchanged
  T kin1 = T(0);
#pragma unroll
  for (int j = 0; j < NET::P; ++j) kin1 = fma_t<T>(p[j], p[j], kin1);
  const T h_prop = -ltp + T(0.5) * kin1;                                                  // :141
  T rate = exp_t<T>(h_cur - h_prop);
  rate = (rate > T(1)) ? T(1) : rate;                                                     // torch.min keeps NaN, :143-146
  if (rate_out) *rate_out = rate;
  return u < rate;                                                                        // :148 (linear space)
}

// HMCDATuner.tune (eeyore/tuners/hmcda_tuner.py:44-59) for one chain, in fp64 like the reference's python floats.
// g = 0.05, t0 = 10, k = 0.75 (:26-30).  `it` = counter.idx + 1.  A NaN rate (diverged trajectory) counts as 0 -- the
// reference would poison its tuner state with NaN for the rest of the run.
struct DaTuner {
  double l, d, m, logeub;
  int has_eub;
};
EB_HD void da_tune(const DaTuner& tn, double rate, long it, bool return_e, double& barh, double& logbare, double& step,
                   int& num_steps) {
  if (!(rate == rate)) rate = 0.0;
  const double d_w = 1.0 / ((double)it + 10.0);
  const double e_w = 1.0 / pow((double)it, 0.75);
  barh = (1.0 - d_w) * barh + d_w * (tn.d - rate);
  double loge = tn.m - sqrt((double)it) * barh / 0.05;
  if (tn.has_eub) loge = loge < tn.logeub ? loge : tn.logeub;
  logbare = e_w * loge + (1.0 - e_w) * logbare;
  step = exp(return_e ? loge : logbare);
  const double q = rint(tn.l / step);                                                     // round half to even, :41-42
  num_steps = q >= 1.0 ? (q < 100000.0 ? (int)q : 100000) : 1;
}

}  // namespace eb
