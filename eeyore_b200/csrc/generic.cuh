// Runtime-shape MLP path: any small fully-connected network the reference's Hyperparameters can describe
// (eeyore/models/mlp.py:9-19,37-50): arbitrary dims, per-layer bias on/off, per-layer activation sigmoid / None.
// It is the fallback behind the compile-time specialisations of mlp_static.cuh: the same arithmetic, but parameters,
// gradient accumulators and activations live in per-thread shared-memory columns (element j of a thread's vector at
// base[j * stride], stride = threads per CTA => bank-conflict free) because their sizes are only known at run time.
// One thread per chain.  Host+device code: tests/hostsim executes it on the CPU against the oracle.
//
// Replaces the same reference lines as mlp_static.cuh / samplers.cuh (model.py:44-55, mlp.py:45-50, stats/loss.py:1-11,
// constants.py:15-18, bayesian_model.py:30-56, log_target_model.py:15-23, metropolis_hastings.py:41-73, mala.py:46-82,
// hmc.py:100-170).
#pragma once
#include "samplers.cuh"
#include "philox.cuh"

namespace eb {

constexpr int kGenMaxLayers = 8;

struct GenNet {
  int nl;                       // number of dense layers
  int dims[kGenMaxLayers + 1];
  int bias[kGenMaxLayers];
  int act[kGenMaxLayers];       // 1 = sigmoid, 0 = identity
  int off[kGenMaxLayers];       // start of layer l in theta (weights, then bias if any)
  int hoff[kGenMaxLayers + 1];  // start of h_l in the activation workspace
  int loss, P, H, maxd;         // H = sum of dims (workspace size), maxd = widest layer
};

inline GenNet make_gen_net(int nl, const int* dims, const int* bias, const int* act, int loss) {
  GenNet n{};
  n.nl = nl; n.loss = loss;
  int o = 0, h = 0, m = 0;
  for (int l = 0; l <= nl; ++l) { n.dims[l] = dims[l]; n.hoff[l] = h; h += dims[l]; m = dims[l] > m ? dims[l] : m; }
  for (int l = 0; l < nl; ++l) {
    n.bias[l] = bias[l]; n.act[l] = act[l]; n.off[l] = o;
    o += (dims[l] + (bias[l] ? 1 : 0)) * dims[l + 1];
  }
  n.P = o; n.H = h; n.maxd = m;
  return n;
}

// per-thread workspace vectors (all strided by the CTA width)
template <typename T> struct GenWork {
  StridedVec<T> h;    // [H] activations of the current row
  StridedVec<T> da;   // [maxd] delta of the layer being processed
  StridedVec<T> db;   // [maxd] delta of the layer below
};

template <typename T, bool GRAD, class TH, class GV>
EB_HD void gen_accumulate_row(const GenNet& n, const TH& th, const T* xr, T y, int cls, const GenWork<T>& w, T& ll, GV& g) {
  for (int i = 0; i < n.dims[0]; ++i) w.h[i] = xr[i];
  for (int l = 0; l < n.nl; ++l) {
    const int din = n.dims[l], dout = n.dims[l + 1], ow = n.off[l], ob = ow + din * dout;
    for (int o = 0; o < dout; ++o) {
      T a = n.bias[l] ? th[ob + o] : T(0);
      for (int i = 0; i < din; ++i) a = fma_t<T>(th[ow + o * din + i], w.h[n.hoff[l] + i], a);
      // the head's sigmoid is applied by the loss below (binary); hidden units apply it here
      const bool head = (l == n.nl - 1);
      w.h[n.hoff[l + 1] + o] = (n.act[l] && !head) ? sigmoid_t<T>(a) : a;
    }
  }
  const int dl = n.dims[n.nl], ho = n.hoff[n.nl];
  // ---- loss and seed (same semantics as head_loss in mlp_static.cuh) ----
  if (n.loss == LOSS_BINARY) {
    const T a0 = w.h[ho];
    T p = sigmoid_t<T>(a0);
    if constexpr (sizeof(T) == 8) { if (a0 < T(-700)) p = sigmoid_ref_tail(a0); }   // as in mlp_static.cuh: head_loss
    T term;
    if (y == T(1) || y == T(0)) {
      const T q = (y == T(1)) ? p : (T(1) - p);
      const T other = (y == T(1)) ? (T(1) - p) : p;
      T lq = head_log_nz<T>(q, q == T(0));
      lq = (q == T(0)) ? -T(INFINITY) : lq;
      term = (other == T(0) || q != q) ? qnan<T>() : lq;
    } else {
      term = log_t<T>(p) * y + log_t<T>(T(1) - p) * (T(1) - y);
    }
    ll += term;
    w.da[0] = (p == T(0) || p == T(1)) ? qnan<T>() : (y - p);
  } else {
    T m = w.h[ho];
    for (int k = 1; k < dl; ++k) m = w.h[ho + k] > m ? w.h[ho + k] : m;
    T s = T(0);
    for (int k = 0; k < dl; ++k) { const T e = exp_nonpos_t<T>(w.h[ho + k] - m); w.da[k] = e; s += e; }
    const T inv = T(1) / s, ls = head_log<T>(s);
    ll += w.h[ho + cls] - m - ls;
    for (int k = 0; k < dl; ++k) w.da[k] = ((k == cls) ? T(1) : T(0)) - w.da[k] * inv;
  }
  // ---- back-propagation ----
  if constexpr (GRAD) {
  for (int l = n.nl - 1; l >= 0; --l) {
    const int din = n.dims[l], dout = n.dims[l + 1], ow = n.off[l], ob = ow + din * dout;
    const StridedVec<T>& dcur = ((n.nl - 1 - l) & 1) ? w.db : w.da;
    const StridedVec<T>& dnext = ((n.nl - 1 - l) & 1) ? w.da : w.db;
    for (int o = 0; o < dout; ++o) {
      const T d = dcur[o];
      for (int i = 0; i < din; ++i) g[ow + o * din + i] = fma_t<T>(d, w.h[n.hoff[l] + i], g[ow + o * din + i]);
      if (n.bias[l]) g[ob + o] += d;
    }
    if (l > 0) {
      for (int i = 0; i < din; ++i) {
        T s = T(0);
        for (int o = 0; o < dout; ++o) s = fma_t<T>(dcur[o], th[ow + o * din + i], s);
        const T hv = w.h[n.hoff[l] + i];
        dnext[i] = n.act[l - 1] ? s * (T(1) - hv) * hv : s;
      }
    }
  }
  }
}

template <typename T, bool GRAD, class TH, class GV>
EB_HD void gen_eval_target(const GenNet& n, const DataView<T>& d, const TH& th, const GenWork<T>& w, T& lt, GV& g,
                           T* ll_out = nullptr, T* lp_out = nullptr) {
  T ll = T(0);
  if constexpr (GRAD) for (int j = 0; j < n.P; ++j) g[j] = T(0);
  for (int i = 0; i < d.n_rows; ++i) {
    T y = T(0);
    int cls = 0;
    if (n.loss == LOSS_BINARY) y = d.y[i]; else cls = d.cls[i];
    gen_accumulate_row<T, GRAD>(n, th, d.x + (size_t)i * n.dims[0], y, cls, w, ll, g);
  }
  T lp = d.lp_const;
  for (int j = 0; j < n.P; ++j) {
    const T dd = th[j] - d.ploc[j];
    lp = fma_t<T>(-(dd * dd), T(0.5) * d.pivar[j], lp);
    if constexpr (GRAD) g[j] = fma_t<T>(-dd, d.pivar[j], g[j]);
  }
  if (d.has_temperature) {
    ll *= d.temperature; lp *= d.temperature;
    if constexpr (GRAD) for (int j = 0; j < n.P; ++j) g[j] *= d.temperature;
  }
  lt = ll + lp;
  if (ll_out) *ll_out = ll;
  if (lp_out) *lp_out = lp;
}

// ---- draws (runtime-P restatements of mh_draw / mala_draw / hmc_draw; z, thp, gp, mom are workspace vectors) ----------
template <typename T>
EB_HD bool gen_mh_draw(const GenNet& n, const DataView<T>& d, const GenWork<T>& w, T prop_scale, bool symmetric,
                       const Cur<T>& cur, T lt_cur, const StridedVec<T>& z, T u, const StridedVec<T>& thp, T& ltp) {
  for (int j = 0; j < n.P; ++j) thp[j] = fma_t<T>(prop_scale, z[j], cur.th[j * cur.stride]);
  int dummy = 0;
  gen_eval_target<T, false>(n, d, thp, w, ltp, dummy);
  T log_rate = ltp - lt_cur;
  if (!symmetric) {
    const T inv2var = T(1) / (T(2) * prop_scale * prop_scale);
    const T lnorm = log_t<T>(prop_scale) + T(kLogSqrt2Pi);
    T lq_f = T(0), lq_b = T(0);
    for (int j = 0; j < n.P; ++j) {
      const T a = thp[j] - cur.th[j * cur.stride], b = cur.th[j * cur.stride] - thp[j];
      lq_f += -(a * a) * inv2var - lnorm;
      lq_b += -(b * b) * inv2var - lnorm;
    }
    log_rate = log_rate - lq_f;
    log_rate = log_rate + lq_b;
  }
  return log_t<T>(u) < log_rate;
}

template <typename T>
EB_HD bool gen_mala_draw(const GenNet& n, const DataView<T>& d, const GenWork<T>& w, T half_step, T sd, const Cur<T>& cur,
                         T lt_cur, const StridedVec<T>& z, T u, const StridedVec<T>& thp, const StridedVec<T>& gp, T& ltp) {
  const T inv2var = T(1) / (T(2) * (sd * sd));
  const T lnorm = log_t<T>(sd) + T(kLogSqrt2Pi);
  T lq_f = T(0);
  for (int j = 0; j < n.P; ++j) {
    const T mean = fma_t<T>(half_step, cur.g[j * cur.stride], cur.th[j * cur.stride]);
    thp[j] = fma_t<T>(sd, z[j], mean);
    const T dd = thp[j] - mean;
    lq_f += -(dd * dd) * inv2var - lnorm;
  }
  gen_eval_target<T, true>(n, d, thp, w, ltp, gp);
  T lq_b = T(0);
  for (int j = 0; j < n.P; ++j) {
    const T mean_p = fma_t<T>(half_step, gp[j], thp[j]);
    const T dd = cur.th[j * cur.stride] - mean_p;
    lq_b += -(dd * dd) * inv2var - lnorm;
  }
  T log_rate = ltp - lt_cur;
  log_rate = log_rate - lq_f;
  log_rate = log_rate + lq_b;
  return log_t<T>(u) < log_rate;
}

// z holds the momentum draw on entry and is updated in place (it is the momentum vector)
template <typename T>
EB_HD bool gen_hmc_draw(const GenNet& n, const DataView<T>& d, const GenWork<T>& w, T eps, T half_eps, int num_steps,
                        const Cur<T>& cur, T lt_cur, const StridedVec<T>& z, T u, const StridedVec<T>& thp,
                        const StridedVec<T>& gp, T& ltp, T* rate_out = nullptr) {
  T kin = T(0);
  for (int j = 0; j < n.P; ++j) kin = fma_t<T>(z[j], z[j], kin);
  const T h_cur = -lt_cur + T(0.5) * kin;
  for (int j = 0; j < n.P; ++j) {
    thp[j] = cur.th[j * cur.stride];
    z[j] = fma_t<T>(half_eps, cur.g[j * cur.stride], z[j]);
  }
  ltp = lt_cur;
  for (int s = 0; s < num_steps; ++s) {
    for (int j = 0; j < n.P; ++j) thp[j] = fma_t<T>(eps, z[j], thp[j]);
    gen_eval_target<T, true>(n, d, thp, w, ltp, gp);
    const T ww = (s == num_steps - 1) ? half_eps : eps;
    for (int j = 0; j < n.P; ++j) z[j] = fma_t<T>(ww, gp[j], z[j]);
  }
  T kin1 = T(0);
  for (int j = 0; j < n.P; ++j) kin1 = fma_t<T>(z[j], z[j], kin1);
  const T h_prop = -ltp + T(0.5) * kin1;
  T rate = exp_t<T>(h_cur - h_prop);
  rate = (rate > T(1)) ? T(1) : rate;
  if (rate_out) *rate_out = rate;
  return u < rate;
}

// standard normals of (chain, iteration) into a workspace vector; same stream layout as philox_normals
template <typename T> EB_HD void gen_philox_normals(const StridedVec<T>& z, int P, RngKey key, uint32_t chain, uint32_t iter) {
  constexpr int PER = sizeof(T) == 8 ? 2 : 4;
  for (int j = 0; j < (P + PER - 1) / PER; ++j) {
    U4 w = philox4x32_10(U4{(uint32_t)j, iter, chain, 0u}, key.k0, key.k1);
    T v[4] = {T(0), T(0), T(0), T(0)};
    if (sizeof(T) == 8) {
      box_muller<T>((T)Uni<double>::from(w.x, w.y), (T)Uni<double>::from(w.z, w.w), &v[0], &v[1]);
    } else {
      box_muller<T>((T)Uni<float>::from(w.x), (T)Uni<float>::from(w.y), &v[0], &v[1]);
      box_muller<T>((T)Uni<float>::from(w.z), (T)Uni<float>::from(w.w), &v[2], &v[3]);
    }
    for (int k = 0; k < PER; ++k)
      if (j * PER + k < P) z[j * PER + k] = v[k];
  }
}

}  // namespace eb
