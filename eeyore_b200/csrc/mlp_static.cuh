// Compile-time MLP: log-likelihood, log-prior, log-target and gradient for one chain, weights in registers.
//
// Replaces (reference paths):
//   eeyore/models/model.py:44-55            flat theta layout: per layer W row-major [out,in], then bias
//   eeyore/models/mlp.py:45-50              forward pass (sigmoid hidden units; sigmoid or identity head)
//   eeyore/stats/loss.py:1-11               naive binary cross-entropy on probabilities, reduction='sum'
//   eeyore/constants/constants.py:15-18     binary / multiclass loss table
//   eeyore/models/bayesian_model.py:30-56   log_lik, log_prior (vector Normal), log_target, temperature
//   eeyore/models/log_target_model.py:15-23 gradient (torch.autograd there; closed-form back-prop here)
//
// Everything is __host__ __device__ so tests/hostsim can execute the very same per-chain code on the CPU
// as a pre-flight check (the shipped library contains only the device instantiations).
#pragma once
#include "common.cuh"
#include <type_traits>

#ifndef EB_ROW_UNROLL
#define EB_ROW_UNROLL 1
#endif
// rows of the data set one lane pushes through the fp64 fast path together (their dependency chains interleave).
// Measured on B200, config 4 (evals/s): batch 1 at 168 registers / 12 warps per SM 14.1e9; batch 2 at 168 registers 13.8e9
// (spills); batch 2 at 250 registers / 8 warps per SM 15.2e9; batch 4 at 254 registers 14.5e9.
#ifndef EB_ROW_BATCH
#define EB_ROW_BATCH 2
#endif

namespace eb {

enum { LOSS_BINARY = 0, LOSS_MULTICLASS = 1 };

// Network with 2 or 3 dense layers (D3 == 0 -> 2 layers), all biases on, sigmoid hidden units;
// head: sigmoid (binary, D_last == 1) or identity (multiclass logits).
template <int LOSS_, int D0_, int D1_, int D2_, int D3_ = 0> struct Net {
  static constexpr int LOSS = LOSS_;
  static constexpr int NL = D3_ > 0 ? 3 : 2;
  static constexpr int D0 = D0_, D1 = D1_, D2 = D2_, D3 = D3_;
  static constexpr int DL = NL == 3 ? D3_ : D2_;
  static constexpr int OFF0 = 0;
  static constexpr int OFF1 = (D0_ + 1) * D1_;
  static constexpr int OFF2 = OFF1 + (D1_ + 1) * D2_;
  static constexpr int P = OFF2 + (NL == 3 ? (D2_ + 1) * D3_ : 0);
  static_assert(LOSS_ != LOSS_BINARY || DL == 1, "binary head has one output");
};

// Per-block view of the (shared-memory resident) data set and prior.
template <typename T> struct DataView {
  const T* x;       // [N, D0]
  const T* y;       // [N] binary targets (LOSS_BINARY)
  const int* cls;   // [N] class index = argmax of the one-hot row (LOSS_MULTICLASS), constants.py:17
  int n_rows;
  bool hard_labels; // every y is exactly 0 or 1 (LOSS_BINARY): the row code is then branch-free
  const T* ploc;    // [P] prior mean
  const T* pivar;   // [P] 1 / scale^2
  T lp_const;       // sum_j ( -log scale_j - log sqrt(2 pi) )
  T temperature;
  bool has_temperature;
};

template <typename T, int DIN, int DOUT, int OFF, bool SIG, class TH>
EB_HD void dense_fwd(const TH& th, const T (&in)[DIN], T (&out)[DOUT]) {
  T pre[DOUT];
#pragma unroll
  for (int o = 0; o < DOUT; ++o) {
    T a = th[OFF + DIN * DOUT + o];
#pragma unroll
    for (int i = 0; i < DIN; ++i) a = fma_t<T>(th[OFF + o * DIN + i], in[i], a);
    pre[o] = a;
  }
  if constexpr (SIG) {
    sigmoid_vec<T, DOUT>(pre, out);
  } else {
#pragma unroll
    for (int o = 0; o < DOUT; ++o) out[o] = pre[o];
  }
}

// fp64 fast path of a sigmoid layer: no saturation / NaN handling inside; `mx` collects the largest |pre-activation|
// high word so that the caller can fall back to the general code (mlp_static.cuh: accumulate_row).
template <typename T, int DIN, int DOUT, int OFF, class TH>
EB_HD void dense_fwd_sig_fast(const TH& th, const T (&in)[DIN], T (&out)[DOUT], int& mx) {
  T pre[DOUT];
#pragma unroll
  for (int o = 0; o < DOUT; ++o) {
    T a = th[OFF + DIN * DOUT + o];
#pragma unroll
    for (int i = 0; i < DIN; ++i) a = fma_t<T>(th[OFF + o * DIN + i], in[i], a);
    pre[o] = a;
  }
#pragma unroll
  for (int o = 0; o < DOUT; ++o) {
    const int ah = abs_hi(pre[o]);
    mx = ah > mx ? ah : mx;
    out[o] = sigmoid_fast(pre[o]);
  }
}

// Accumulates dW += delta (x) in, db += delta; if PROP also back-propagates through the sigmoid that produced `in`
// (autograd order: grad * (1 - out) * out).
template <typename T, int DIN, int DOUT, int OFF, bool PROP, class TH, class GV>
EB_HD void dense_bwd(const TH& th, const T (&in)[DIN], const T (&dout)[DOUT], GV& g, T (&din)[DIN]) {
#pragma unroll
  for (int o = 0; o < DOUT; ++o) {
#pragma unroll
    for (int i = 0; i < DIN; ++i) g[OFF + o * DIN + i] = fma_t<T>(dout[o], in[i], g[OFF + o * DIN + i]);
    g[OFF + DIN * DOUT + o] += dout[o];
  }
  if (PROP) {
#pragma unroll
    for (int i = 0; i < DIN; ++i) {
      T s = T(0);
#pragma unroll
      for (int o = 0; o < DOUT; ++o) s = fma_t<T>(dout[o], th[OFF + o * DIN + i], s);
      din[i] = s * fma_t<T>(-in[i], in[i], in[i]);   // out (1 - out) in one FMA
    }
  }
}

template <typename T> EB_HD T rcp_sum_t(T s) { return T(1) / s; }
template <> EB_HD double rcp_sum_t<double>(double s) { return rcp_ge1(s); }
// the fp32 fast path (HARD = true marks it for the multiclass head): MUFU-based exp / reciprocal / log
template <typename T, bool FAST> EB_HD T head_exp(T a) { return exp_nonpos_t<T>(a); }
template <> EB_HD float head_exp<float, true>(float a) { return exp_nonpos_fast(a); }
template <typename T, bool FAST> EB_HD T head_rcp(T s) { return rcp_sum_t<T>(s); }
template <> EB_HD float head_rcp<float, true>(float s) { return rcp_fast(s); }
template <typename T, bool FAST> EB_HD T head_logsum(T s);
template <> EB_HD double head_logsum<double, true>(double s) { return log_pos_normal(s > 0.0 ? s : 1.0); }
template <> EB_HD double head_logsum<double, false>(double s) { return log_pos_normal(s > 0.0 ? s : 1.0); }
template <> EB_HD float head_logsum<float, true>(float s) { return log_fast(s); }
template <> EB_HD float head_logsum<float, false>(float s) { return logf(s); }
template <typename T> EB_HD T head_log(T q) { return log_t<T>(q); }
// fp64: q is a probability in [0, 1]; 0 is patched by the caller, tiny values are normal numbers (>= 1e-304)
template <> EB_HD double head_log<double>(double q) { return log_pos_normal(q > 0.0 ? q : 1.0); }
// the same with the zero test supplied by the caller
template <typename T> EB_HD T head_log_nz(T q, bool q_zero) { return head_log<T>(q); }
template <> EB_HD double head_log_nz<double>(double q, bool q_zero) { return log_prob_any(q_zero ? 1.0 : q); }

// Head: returns this row's log-likelihood term and the seed d ll / d a_L.
template <typename T, class NET, bool HARD = false>
EB_HD T head_loss(T (&a)[NET::DL], T y, int cls, T (&delta)[NET::DL], T* p_out) {
  if constexpr (NET::LOSS == LOSS_BINARY) {
    T p = sigmoid_t<T>(a[0]);
    // the reference's p is subnormal below a = -708 and exactly 0 once exp(-a) overflows (a < -709.78); the fp64 fast
    // sigmoid saturates to 0 from -708 on, so this (general, rarely executed) path evaluates the tail like the reference
    if constexpr (sizeof(T) == 8) { if (a[0] < T(-700)) p = sigmoid_ref_tail(a[0]); }
    if (p_out) *p_out = p;
    T term;
    // loss.py:2 evaluates log(p)*y + log(1-p)*(1-y); for y in {0,1} one product is 0 * log(.), which is NaN
    // exactly when that log is -inf -- reproduced without evaluating the second log and without
    // branching (one log of the selected argument; selects restore the special cases).
    const bool y1 = prob_is_one<T>(y), y0 = prob_is_zero<T>(y);
    const bool p0 = prob_is_zero<T>(p), p1 = prob_is_one<T>(p);
    if (HARD || y1 || y0) {
      const T q = y1 ? p : (T(1) - p);                   // the probability of the observed label
      const bool q_zero = y1 ? p0 : p1;                  // 1 - p == 0 exactly when p == 1
      const bool other_zero = y1 ? p1 : p0;              // the complement: 0 there means 0 * log(0) = NaN
      const T lq = head_log_nz<T>(q, q_zero);
      term = (other_zero || prob_is_nan<T>(p)) ? qnan<T>() : (q_zero ? -T(INFINITY) : lq);
    } else {
      term = log_t<T>(p) * y + log_t<T>(T(1) - p) * (T(1) - y);
    }
    // autograd of the naive form gives (y/p - (1-y)/(1-p)) (1-p) p = y - p, and NaN when p hits 0 or 1
    delta[0] = (p0 || p1) ? qnan<T>() : (y - p);
    return term;
  } else {
    constexpr int K = NET::DL;
    T m = a[0];
#pragma unroll
    for (int k = 1; k < K; ++k) m = a[k] > m ? a[k] : m;
    T e[K];
    T s = T(0);
#pragma unroll
    for (int k = 0; k < K; ++k) { e[k] = head_exp<T, HARD>(a[k] - m); s += e[k]; }
    const T inv = head_rcp<T, HARD>(s);   // s >= 1 (the maximal logit contributes e^0)
    const T ls = head_logsum<T, HARD>(s);
    T term = T(0);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const bool hit = (k == cls);
      if (hit) term = a[k] - m - ls;
      delta[k] = (hit ? T(1) : T(0)) - e[k] * inv;
    }
    return term;
  }
}

// One data row: forward, loss, (optionally) backward; accumulates into ll and g.
// fp64 with hard labels takes a fast forward pass first: sigmoids without saturation selects and, for the binary head, no
// special-case selects either -- valid while every hidden pre-activation is below 708 in magnitude and the head's below 36
// (then 0 < p < 1 strictly and nothing is NaN); otherwise the lane redoes the forward pass with the general code, which
// reproduces the reference's saturation / NaN semantics.  Both paths evaluate identical arithmetic where both are valid.
// (The FP64 pipe and the dispatch port are the bound: the selects were a fifth of the instructions of a row.)
// VALUE = false (the inner leapfrog steps of an HMC trajectory, which use the gradient only): the fp64 fast path skips the
// logarithm of the row's term; everything that decides between the fast and the general path is unchanged.
template <typename T, class NET, bool GRAD, bool HARD = false, bool VALUE = true, class TH, class GV>
EB_HD void accumulate_row(const TH& th, const T* xr, T y, int cls, T& ll, GV& g) {
  T h0[NET::D0];
#pragma unroll
  for (int i = 0; i < NET::D0; ++i) h0[i] = xr[i];
  T h1[NET::D1];
  T h2[NET::NL == 3 ? NET::D2 : 1];
  T dl[NET::DL];
  T term = T(0);
  bool done = false;
  if constexpr (HARD || NET::LOSS != LOSS_BINARY) {
    int mx = 0;
    T a[NET::DL];
    dense_fwd_sig_fast<T, NET::D0, NET::D1, NET::OFF0>(th, h0, h1, mx);
    if constexpr (NET::NL == 2) {
      dense_fwd<T, NET::D1, NET::D2, NET::OFF1, false>(th, h1, a);
    } else {
      dense_fwd_sig_fast<T, NET::D1, NET::D2, NET::OFF1>(th, h1, h2, mx);
      dense_fwd<T, NET::D2, NET::DL, NET::OFF2, false>(th, h2, a);
    }
    if constexpr (NET::LOSS == LOSS_BINARY) {
      const T p = sigmoid_fast(a[0]);
      const bool y1 = prob_is_one<T>(y);
      if constexpr (VALUE) term = log_prob_fast(y1 ? p : T(1) - p);
      dl[0] = y - p;
      done = mx <= FastBounds<T>::hidden && abs_hi(a[0]) <= FastBounds<T>::head;
    } else {
      term = head_loss<T, NET, true>(a, y, cls, dl, (T*)nullptr);
      done = mx <= FastBounds<T>::hidden;
    }
  }
  if (!done) {
    dense_fwd<T, NET::D0, NET::D1, NET::OFF0, true>(th, h0, h1);
    T a[NET::DL];
    if constexpr (NET::NL == 2) {
      dense_fwd<T, NET::D1, NET::D2, NET::OFF1, false>(th, h1, a);
    } else {
      dense_fwd<T, NET::D1, NET::D2, NET::OFF1, true>(th, h1, h2);
      dense_fwd<T, NET::D2, NET::DL, NET::OFF2, false>(th, h2, a);
    }
    term = head_loss<T, NET, HARD && NET::LOSS == LOSS_BINARY>(a, y, cls, dl, (T*)nullptr);
  }
  ll += term;
  if constexpr (GRAD) {
    T d1[NET::D1], d0[NET::D0];
    if constexpr (NET::NL == 2) {
      dense_bwd<T, NET::D1, NET::D2, NET::OFF1, true>(th, h1, dl, g, d1);
    } else {
      T d2[NET::D2];
      dense_bwd<T, NET::D2, NET::DL, NET::OFF2, true>(th, h2, dl, g, d2);
      dense_bwd<T, NET::D1, NET::D2, NET::OFF1, true>(th, h1, d2, g, d1);
    }
    dense_bwd<T, NET::D0, NET::D1, NET::OFF0, false>(th, h0, d1, g, d0);
  }
}

// R rows at once through the fp64 fast path (see accumulate_row), every stage written across the rows.  Returns false --
// with nothing accumulated -- when a bound of the fast path is violated; the caller then takes the rows one by one.
template <typename T, int DIN, int DOUT, int OFF, int R, class TH>
EB_HD void layer_fast_rows(const TH& th, const T (&in)[R][DIN], T (&out)[R][DOUT], int& mx) {
  T pre[R * DOUT], s[R * DOUT];
#pragma unroll
  for (int o = 0; o < DOUT; ++o) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      T a = th[OFF + DIN * DOUT + o];
#pragma unroll
      for (int i = 0; i < DIN; ++i) a = fma_t<T>(th[OFF + o * DIN + i], in[r][i], a);
      pre[r * DOUT + o] = a;
    }
  }
  sigmoid_fast_vec<R * DOUT>(pre, s, mx);
#pragma unroll
  for (int r = 0; r < R; ++r) {
#pragma unroll
    for (int o = 0; o < DOUT; ++o) out[r][o] = s[r * DOUT + o];
  }
}

template <typename T, class NET, bool GRAD, int R, bool VALUE = true, class TH, class GV>
EB_HD bool accumulate_rows_fast(const TH& th, const T* const (&xr)[R], const T (&y)[R], const int (&cls)[R], T& ll, GV& g) {
  {
  T h0[R][NET::D0], h1[R][NET::D1], h2[R][NET::NL == 3 ? NET::D2 : 1], a[R][NET::DL], dl[R][NET::DL], term[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
#pragma unroll
    for (int i = 0; i < NET::D0; ++i) h0[r][i] = xr[r][i];
  }
  int mx = 0;
  layer_fast_rows<T, NET::D0, NET::D1, NET::OFF0, R>(th, h0, h1, mx);
  if constexpr (NET::NL == 3) layer_fast_rows<T, NET::D1, NET::D2, NET::OFF1, R>(th, h1, h2, mx);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if constexpr (NET::NL == 2) dense_fwd<T, NET::D1, NET::D2, NET::OFF1, false>(th, h1[r], a[r]);
    else dense_fwd<T, NET::D2, NET::DL, NET::OFF2, false>(th, h2[r], a[r]);
  }
  bool ok = true;
  if constexpr (NET::LOSS == LOSS_BINARY) {
    T a0[R], p[R];
    int mh = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) a0[r] = a[r][0];
    sigmoid_fast_vec<R>(a0, p, mh);
    ok = mh <= FastBounds<T>::head;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      term[r] = T(0);
      if constexpr (VALUE) term[r] = log_prob_fast(prob_is_one<T>(y[r]) ? p[r] : T(1) - p[r]);
      dl[r][0] = y[r] - p[r];
    }
  } else {
#pragma unroll
    for (int r = 0; r < R; ++r) term[r] = head_loss<T, NET, true>(a[r], y[r], cls[r], dl[r], (T*)nullptr);
  }
  if (!(ok && mx <= FastBounds<T>::hidden)) return false;
  if constexpr (VALUE) {
#pragma unroll
    for (int r = 0; r < R; ++r) ll += term[r];
  }
  if constexpr (GRAD) {
    T d1[R][NET::D1], d0[NET::D0];
    if constexpr (NET::NL == 2) {
#pragma unroll
      for (int r = 0; r < R; ++r) dense_bwd<T, NET::D1, NET::D2, NET::OFF1, true>(th, h1[r], dl[r], g, d1[r]);
    } else {
      T d2[R][NET::D2];
#pragma unroll
      for (int r = 0; r < R; ++r) dense_bwd<T, NET::D2, NET::DL, NET::OFF2, true>(th, h2[r], dl[r], g, d2[r]);
#pragma unroll
      for (int r = 0; r < R; ++r) dense_bwd<T, NET::D1, NET::D2, NET::OFF1, true>(th, h1[r], d2[r], g, d1[r]);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) dense_bwd<T, NET::D0, NET::D1, NET::OFF0, false>(th, h0[r], d1[r], g, d0);
  }
  return true;
  }
}

// Network output for one row (MLP.forward): probability (binary) or logits (multiclass).
template <typename T, class NET, class TH> EB_HD void forward_row(const TH& th, const T* xr, T (&out)[NET::DL]) {
  T h0[NET::D0];
#pragma unroll
  for (int i = 0; i < NET::D0; ++i) h0[i] = xr[i];
  T h1[NET::D1];
  dense_fwd<T, NET::D0, NET::D1, NET::OFF0, true>(th, h0, h1);
  constexpr bool sig_head = NET::LOSS == LOSS_BINARY;
  if constexpr (NET::NL == 2) {
    dense_fwd<T, NET::D1, NET::D2, NET::OFF1, sig_head>(th, h1, out);
  } else {
    T h2[NET::D2];
    dense_fwd<T, NET::D1, NET::D2, NET::OFF1, true>(th, h1, h2);
    dense_fwd<T, NET::D2, NET::DL, NET::OFF2, sig_head>(th, h2, out);
  }
}

// log_target (and gradient) of one chain.  The G lanes of a chain group split the data rows (lane `sub` takes rows
// sub, sub+G, ...), then all-reduce; every lane returns the same lt / g.  bayesian_model.py:52-56.
// VALUE = false: gradient only -- lt is left untouched, the row terms skip their logarithm on the fp64 fast path and the
// prior contributes its gradient alone (one FMA per parameter).  HMC's inner leapfrog steps need nothing else
// (eeyore/samplers/hmc.py:110-119 evaluates the target there as well, but only the last value reaches the accept test).
template <typename T, class NET, int G, bool GRAD, bool VALUE = true, class TH, class GV>
EB_HD void eval_target(const DataView<T>& d, int sub, const TH& th, T& lt, GV& g, T* ll_out = nullptr,
                       T* lp_out = nullptr) {
  T ll = T(0);
  if constexpr (GRAD) {
#pragma unroll
    for (int j = 0; j < NET::P; ++j) g[j] = T(0);
  }
  // HARD: all labels are exactly 0 / 1 (uniform over the block, found once when the data set is staged), so the row code
  // has no soft-label branch and is one basic block: with two rows per trip the instruction scheduler interleaves their
  // dependency chains (the per-row critical path -- three sigmoid layers and a log in sequence -- is latency-bound).
  auto rows = [&](auto hard_tag) {
    constexpr bool HARD = decltype(hard_tag)::value;
    auto one_row = [&](int i) {
      T y = T(0);
      int cls = 0;
      if constexpr (NET::LOSS == LOSS_BINARY) y = d.y[i]; else cls = d.cls[i];
      accumulate_row<T, NET, GRAD, HARD, VALUE>(th, d.x + i * NET::D0, y, cls, ll, g);
    };
    int i = sub;
#if EB_ROW_BATCH >= 2
    if constexpr (HARD || NET::LOSS != LOSS_BINARY) {
#ifdef EB_ROW_BATCH_GRADONLY
      constexpr int R = VALUE ? EB_ROW_BATCH : EB_ROW_BATCH_GRADONLY;
#else
      constexpr int R = EB_ROW_BATCH;
#endif
      for (; i + (R - 1) * G < d.n_rows; i += R * G) {
        const T* xr[R];
        T yr[R];
        int cr[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          xr[r] = d.x + (i + r * G) * NET::D0;
          yr[r] = T(0); cr[r] = 0;
          if constexpr (NET::LOSS == LOSS_BINARY) yr[r] = d.y[i + r * G]; else cr[r] = d.cls[i + r * G];
        }
        if (!accumulate_rows_fast<T, NET, GRAD, R, VALUE>(th, xr, yr, cr, ll, g)) {
#pragma unroll 1
          for (int r = 0; r < R; ++r) one_row(i + r * G);
        }
      }
    }
#endif
#if EB_ROW_UNROLL >= 2
    for (; i + G < d.n_rows; i += 2 * G) {
      one_row(i);
      one_row(i + G);
    }
#endif
    for (; i < d.n_rows; i += G) one_row(i);
  };
  if (NET::LOSS != LOSS_BINARY || d.hard_labels) rows(std::true_type{}); else rows(std::false_type{});
#if defined(__CUDA_ARCH__)
  if constexpr (G > 1) {
    if constexpr (VALUE) ll = group_allreduce<G>(ll);
    if constexpr (GRAD) {
#pragma unroll
      for (int j = 0; j < NET::P; ++j) g[j] = group_allreduce<G>(g[j]);
    }
  }
#endif
  // vector Normal prior: sum_j -(theta_j - loc_j)^2 / (2 scale_j^2) - log scale_j - log sqrt(2 pi); bayesian_model.py:46-50
  // (initialising the accumulators with the prior gradient instead would save an add per parameter, but keeps all of them
  // live through the peeled first row: measured as spills)
  if constexpr (!VALUE) {
    static_assert(GRAD, "a gradient-only evaluation needs GRAD");
#pragma unroll
    for (int j = 0; j < NET::P; ++j) {
      g[j] = fma_t<T>(d.ploc[j] - th[j], d.pivar[j], g[j]);
      if (d.has_temperature) g[j] *= d.temperature;
    }
    return;
  }
  T qp[4] = {T(0), T(0), T(0), T(0)};   // four partial sums: one chain of P dependent FMAs would be pure latency
#pragma unroll
  for (int j = 0; j < NET::P; ++j) {
    const T dd = th[j] - d.ploc[j];
    const T w = dd * d.pivar[j];
    qp[j & 3] = fma_t<T>(dd, w, qp[j & 3]);
    if constexpr (GRAD) g[j] -= w;
  }
  const T qs = (qp[0] + qp[1]) + (qp[2] + qp[3]);
  const T lp_raw = fma_t<T>(T(-0.5), qs, d.lp_const);
  T lp = lp_raw;
  if (d.has_temperature) {  // both terms scaled, bayesian_model.py:33-34,48-49
    ll *= d.temperature; lp *= d.temperature;
    if constexpr (GRAD) {
#pragma unroll
      for (int j = 0; j < NET::P; ++j) g[j] *= d.temperature;
    }
  }
  lt = ll + lp;
  if (ll_out) *ll_out = ll;
  if (lp_out) *lp_out = lp;
}

}  // namespace eb
