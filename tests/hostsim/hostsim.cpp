// TEST INFRASTRUCTURE ONLY -- never linked into libeeyore_b200.so and never used by the product path.
//
// Compiles the per-chain __host__ __device__ code of eeyore_b200/csrc (mlp_static.cuh, samplers.cuh, philox.cuh)
// for the CPU with g++ so that the CPU-only test-suite can execute exactly the arithmetic the CUDA kernels run
// (one lane per chain) and compare it with the oracle and the golden vectors before any GPU time is spent.
// The thread mapping, shared-memory staging and shuffles of chain_kernels.cuh are NOT exercised here; the
// `-m gpu` tests cover those through the real C ABI.
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../eeyore_b200/csrc/samplers.cuh"
#include "../../eeyore_b200/csrc/philox.cuh"
#include "../../eeyore_b200/csrc/generic.cuh"
#include "../../include/eeyore_b200.h"

using namespace eb;

template <typename T, class NET>
static DataView<T> make_view(const T* x, const T* y, int N, const T* loc, const T* scale, int has_t, double temp,
                             std::vector<T>& ys, std::vector<int>& cls, std::vector<T>& pivar) {
  ys.assign(N, T(0)); cls.assign(N, 0); pivar.resize(NET::P);
  if (NET::LOSS == LOSS_BINARY) for (int i = 0; i < N; ++i) ys[i] = y[i];
  else for (int i = 0; i < N; ++i) {
    int best = 0; T bv = y[(size_t)i * NET::DL];
    for (int k = 1; k < NET::DL; ++k) if (y[(size_t)i * NET::DL + k] > bv) { bv = y[(size_t)i * NET::DL + k]; best = k; }
    cls[i] = best;
  }
  T c = T(0);
  for (int j = 0; j < NET::P; ++j) { pivar[j] = T(1) / (scale[j] * scale[j]); c += -log_t<T>(scale[j]) - T(kLogSqrt2Pi); }
  DataView<T> d;
  d.hard_labels = true;
  for (int i = 0; i < N; ++i) d.hard_labels = d.hard_labels && (ys[i] == T(0) || ys[i] == T(1));
  d.x = x; d.y = ys.data(); d.cls = cls.data(); d.n_rows = N; d.ploc = loc; d.pivar = pivar.data(); d.lp_const = c;
  d.temperature = (T)temp; d.has_temperature = has_t != 0;
  return d;
}

template <typename T, class NET>
static void eval_all(int64_t C, const T* theta, const T* x, const T* y, int N, const T* loc, const T* scale, int has_t,
                     double temp, T* out_lt, T* out_g) {
  std::vector<T> ys, pivar; std::vector<int> cls;
  DataView<T> d = make_view<T, NET>(x, y, N, loc, scale, has_t, temp, ys, cls, pivar);
  for (int64_t c = 0; c < C; ++c) {
    T th[NET::P], g[NET::P], lt;
    for (int j = 0; j < NET::P; ++j) th[j] = theta[c * NET::P + j];
    if (out_g) { eval_target<T, NET, 1, true>(d, 0, th, lt, g); for (int j = 0; j < NET::P; ++j) out_g[c * NET::P + j] = g[j]; }
    else { int dummy = 0; eval_target<T, NET, 1, false>(d, 0, th, lt, dummy); }
    out_lt[c] = lt;
  }
}

template <typename T, class NET> static void run_all(int kind, const eeyore_b200_run_params& p) {
  constexpr int P = NET::P;
  std::vector<T> ys, pivar; std::vector<int> cls;
  DataView<T> d = make_view<T, NET>((const T*)p.x, (const T*)p.y, (int)p.n_rows, (const T*)p.prior_loc,
                                    (const T*)p.prior_scale, p.has_temperature, p.temperature, ys, cls, pivar);
  T* theta = (T*)p.theta; T* target = (T*)p.target; T* grad = (T*)p.grad;
  const T step = (T)p.step, half_step = T(0.5) * step, sd = sqrt_t<T>(step);
  RngKey key{(uint32_t)(p.seed & 0xffffffffu), (uint32_t)(p.seed >> 32)};
  const int64_t thin = p.thin < 1 ? 1 : p.thin;
  const int64_t sc = (p.st_chain == 0 && p.st_param == 0) ? P : p.st_chain;
  const int64_t sp = (p.st_chain == 0 && p.st_param == 0) ? 1 : p.st_param;
  for (int64_t c = 0; c < p.n_chains; ++c) {
    Cur<T> cur{theta + c * sc, grad ? grad + c * sc : nullptr, (long)sp};
    T mom[P];
    T lt_cur = target[c];
    uint32_t nacc = 0;
    const uint32_t gchain = (uint32_t)p.chain_offset + (uint32_t)c;
    const bool tuned = kind == 2 && p.tuner_state != nullptr;
    DaTuner tn{p.tuner_l, p.tuner_d, p.tuner_m, p.tuner_logeub, p.tuner_has_eub};
    double tn_barh = 0, tn_logbare = 0, tn_step = 0;
    int num_steps = p.num_steps;
    T cstep = step, chalf = half_step;
    if (tuned) {
      tn_barh = p.tuner_state[c]; tn_logbare = p.tuner_state[p.n_chains + c]; tn_step = p.tuner_state[2 * p.n_chains + c];
      num_steps = (int)p.tuner_state[3 * p.n_chains + c];
      cstep = (T)tn_step; chalf = (T)(0.5 * tn_step);
    }
    for (int64_t t = 0; t < p.n_iters; ++t) {
      T z[P], thp[P], gp[P], u, ltp;
      if (p.rng_mode == EEYORE_B200_RNG_PHILOX) {
        philox_normals<T, P>(z, key, gchain, (uint32_t)p.iter_offset + (uint32_t)t);
        u = philox_uniform<T>(key, gchain, (uint32_t)p.iter_offset + (uint32_t)t);
      } else {
        const T* zt = (const T*)p.z_tape + ((size_t)t * p.n_chains + c) * P;
        for (int j = 0; j < P; ++j) z[j] = zt[j];
        u = ((const T*)p.u_tape)[(size_t)t * p.n_chains + c];
      }
      bool acc;
      if (kind == 0) acc = mh_draw<T, NET, 1>(d, 0, step, p.symmetric != 0, cur, lt_cur, z, u, thp, ltp);
      else if (kind == 1) acc = mala_draw<T, NET, 1>(d, 0, half_step, sd, cur, lt_cur, z, u, thp, gp, ltp);
      else {
        T rate;
        { StridedVec<T> pm{mom, 1}; acc = hmc_draw<T, NET, 1>(d, 0, cstep, chalf, num_steps, cur, lt_cur, z, pm, u, thp, gp, ltp, &rate); }
        if (tuned && t < p.tuner_burnin) {
          da_tune(tn, (double)rate, p.tuner_iter0 + t + 1, t != p.tuner_burnin - 1, tn_barh, tn_logbare, tn_step, num_steps);
          cstep = (T)tn_step; chalf = (T)(0.5 * tn_step);
        }
      }
      if (acc) { lt_cur = ltp; ++nacc; for (int j = 0; j < P; ++j) { cur.th[j * sp] = thp[j]; if (kind != 0) cur.g[j * sp] = gp[j]; } }
      if (t >= p.n_burnin && (t - p.n_burnin) % thin == 0) {
        const int64_t s = (t - p.n_burnin) / thin;
        if (p.out_samples) for (int j = 0; j < P; ++j) ((T*)p.out_samples)[s * p.ss_iter + c * p.ss_chain + j * p.ss_param] = cur.th[j * sp];
        if (p.out_grad && kind != 0) for (int j = 0; j < P; ++j) ((T*)p.out_grad)[s * p.ss_iter + c * p.ss_chain + j * p.ss_param] = cur.g[j * sp];
        if (p.out_target) ((T*)p.out_target)[s * p.n_chains + c] = lt_cur;
        if (p.out_accepted) p.out_accepted[s * p.n_chains + c] = acc ? 1 : 0;
      }
    }
    target[c] = lt_cur;
    if (p.accept_count) p.accept_count[c] += nacc;
    if (tuned) {
      p.tuner_state[c] = tn_barh; p.tuner_state[p.n_chains + c] = tn_logbare; p.tuner_state[2 * p.n_chains + c] = tn_step;
      p.tuner_state[3 * p.n_chains + c] = (double)num_steps;
    }
  }
}

#define NETS(X) X(221, LOSS_BINARY, 2, 2, 1) X(2321, LOSS_BINARY, 2, 3, 2, 1) X(433, LOSS_MULTICLASS, 4, 3, 3) X(4323, LOSS_MULTICLASS, 4, 3, 2, 3)

extern "C" {

int hostsim_eval(int arch, int dtype, int64_t C, const void* theta, const void* x, const void* y, int64_t N,
                 const void* loc, const void* scale, int has_t, double temp, void* out_lt, void* out_g) {
#define X(NAME, ...)                                                                                                   \
  if (arch == NAME) {                                                                                                  \
    using NET = Net<__VA_ARGS__>;                                                                                      \
    if (dtype == EEYORE_B200_F64) eval_all<double, NET>(C, (const double*)theta, (const double*)x, (const double*)y, (int)N, (const double*)loc, (const double*)scale, has_t, temp, (double*)out_lt, (double*)out_g); \
    else eval_all<float, NET>(C, (const float*)theta, (const float*)x, (const float*)y, (int)N, (const float*)loc, (const float*)scale, has_t, temp, (float*)out_lt, (float*)out_g); \
    return 0;                                                                                                          \
  }
  NETS(X)
#undef X
  return -1;
}

int hostsim_run(int kind, int arch, int dtype, const eeyore_b200_run_params* p) {
#define X(NAME, ...)                                                  \
  if (arch == NAME) {                                                 \
    using NET = Net<__VA_ARGS__>;                                     \
    if (dtype == EEYORE_B200_F64) run_all<double, NET>(kind, *p);     \
    else run_all<float, NET>(kind, *p);                               \
    return 0;                                                         \
  }
  NETS(X)
#undef X
  return -1;
}

int hostsim_philox(int dtype, int64_t C, int P, uint64_t seed, uint64_t iter, uint64_t chain0, void* z, void* u) {
  RngKey key{(uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32)};
  for (int64_t c = 0; c < C; ++c) {
    uint32_t gc = (uint32_t)chain0 + (uint32_t)c;
    if (dtype == EEYORE_B200_F64) {
      for (int j = 0; j < (P + 1) / 2; ++j) {
        U4 w = philox4x32_10(U4{(uint32_t)j, (uint32_t)iter, gc, 0u}, key.k0, key.k1);
        double a, b; box_muller<double>(Uni<double>::from(w.x, w.y), Uni<double>::from(w.z, w.w), &a, &b);
        ((double*)z)[c * P + 2 * j] = a; if (2 * j + 1 < P) ((double*)z)[c * P + 2 * j + 1] = b;
      }
      ((double*)u)[c] = philox_uniform<double>(key, gc, (uint32_t)iter);
    } else {
      for (int j = 0; j < (P + 3) / 4; ++j) {
        U4 w = philox4x32_10(U4{(uint32_t)j, (uint32_t)iter, gc, 0u}, key.k0, key.k1);
        float v[4]; box_muller<float>(Uni<float>::from(w.x), Uni<float>::from(w.y), &v[0], &v[1]);
        box_muller<float>(Uni<float>::from(w.z), Uni<float>::from(w.w), &v[2], &v[3]);
        for (int k = 0; k < 4; ++k) if (4 * j + k < P) ((float*)z)[c * P + 4 * j + k] = v[k];
      }
      ((float*)u)[c] = philox_uniform<float>(key, gc, (uint32_t)iter);
    }
  }
  return 0;
}
}

// ---- runtime-shape path (generic.cuh) on the CPU: per-thread vectors are plain arrays (stride 1) ---------------------
template <typename T>
static DataView<T> gen_view(const GenNet& n, const T* x, const T* y, int N, const T* loc, const T* scale, int has_t, double temp,
                            std::vector<T>& ys, std::vector<int>& cls, std::vector<T>& pivar) {
  const int dl = n.dims[n.nl];
  ys.assign(N, T(0)); cls.assign(N, 0); pivar.resize(n.P);
  if (n.loss == LOSS_BINARY) for (int i = 0; i < N; ++i) ys[i] = y[i];
  else for (int i = 0; i < N; ++i) {
    int best = 0; T bv = y[(size_t)i * dl];
    for (int k = 1; k < dl; ++k) if (y[(size_t)i * dl + k] > bv) { bv = y[(size_t)i * dl + k]; best = k; }
    cls[i] = best;
  }
  T c = T(0);
  for (int j = 0; j < n.P; ++j) { pivar[j] = T(1) / (scale[j] * scale[j]); c += -log_t<T>(scale[j]) - T(kLogSqrt2Pi); }
  DataView<T> d;
  d.hard_labels = true;
  for (int i = 0; i < N; ++i) d.hard_labels = d.hard_labels && (ys[i] == T(0) || ys[i] == T(1));
  d.x = x; d.y = ys.data(); d.cls = cls.data(); d.n_rows = N; d.ploc = loc; d.pivar = pivar.data(); d.lp_const = c;
  d.temperature = (T)temp; d.has_temperature = has_t != 0;
  return d;
}

template <typename T>
static void gen_eval_all(const GenNet& n, int64_t C, const T* theta, const T* x, const T* y, int N, const T* loc, const T* scale,
                         int has_t, double temp, T* out_lt, T* out_g) {
  std::vector<T> ys, pivar; std::vector<int> cls;
  DataView<T> d = gen_view<T>(n, x, y, N, loc, scale, has_t, temp, ys, cls, pivar);
  std::vector<T> hbuf(n.H), da(n.maxd), db(n.maxd), th(n.P), g(n.P);
  GenWork<T> w{StridedVec<T>{hbuf.data(), 1}, StridedVec<T>{da.data(), 1}, StridedVec<T>{db.data(), 1}};
  for (int64_t c = 0; c < C; ++c) {
    for (int j = 0; j < n.P; ++j) th[j] = theta[c * n.P + j];
    StridedVec<T> thv{th.data(), 1}, gv{g.data(), 1};
    T lt;
    if (out_g) { gen_eval_target<T, true>(n, d, thv, w, lt, gv); for (int j = 0; j < n.P; ++j) out_g[c * n.P + j] = g[j]; }
    else { int dummy = 0; gen_eval_target<T, false>(n, d, thv, w, lt, dummy); }
    out_lt[c] = lt;
  }
}

template <typename T> static void gen_run_all(const GenNet& n, int kind, const eeyore_b200_run_params& p) {
  const int P = n.P;
  std::vector<T> ys, pivar; std::vector<int> cls;
  DataView<T> d = gen_view<T>(n, (const T*)p.x, (const T*)p.y, (int)p.n_rows, (const T*)p.prior_loc, (const T*)p.prior_scale,
                              p.has_temperature, p.temperature, ys, cls, pivar);
  T* theta = (T*)p.theta; T* target = (T*)p.target; T* grad = (T*)p.grad;
  const T step = (T)p.step, half_step = T(0.5) * step, sd = sqrt_t<T>(step);
  RngKey key{(uint32_t)(p.seed & 0xffffffffu), (uint32_t)(p.seed >> 32)};
  const int64_t thin = p.thin < 1 ? 1 : p.thin;
  std::vector<T> hbuf(n.H), da(n.maxd), db(n.maxd), zb(P), thb(P), gb(P);
  GenWork<T> w{StridedVec<T>{hbuf.data(), 1}, StridedVec<T>{da.data(), 1}, StridedVec<T>{db.data(), 1}};
  StridedVec<T> z{zb.data(), 1}, thp{thb.data(), 1}, gp{gb.data(), 1};
  for (int64_t c = 0; c < p.n_chains; ++c) {
    Cur<T> cur{theta + c * P, grad ? grad + c * P : nullptr, 1};
    T lt_cur = target[c];
    uint32_t nacc = 0;
    const uint32_t gchain = (uint32_t)p.chain_offset + (uint32_t)c;
    for (int64_t t = 0; t < p.n_iters; ++t) {
      T u, ltp;
      if (p.rng_mode == EEYORE_B200_RNG_PHILOX) {
        gen_philox_normals<T>(z, P, key, gchain, (uint32_t)p.iter_offset + (uint32_t)t);
        u = philox_uniform<T>(key, gchain, (uint32_t)p.iter_offset + (uint32_t)t);
      } else {
        const T* zt = (const T*)p.z_tape + ((size_t)t * p.n_chains + c) * P;
        for (int j = 0; j < P; ++j) z[j] = zt[j];
        u = ((const T*)p.u_tape)[(size_t)t * p.n_chains + c];
      }
      bool acc;
      if (kind == 0) acc = gen_mh_draw<T>(n, d, w, step, p.symmetric != 0, cur, lt_cur, z, u, thp, ltp);
      else if (kind == 1) acc = gen_mala_draw<T>(n, d, w, half_step, sd, cur, lt_cur, z, u, thp, gp, ltp);
      else acc = gen_hmc_draw<T>(n, d, w, step, half_step, p.num_steps, cur, lt_cur, z, u, thp, gp, ltp);
      if (acc) { lt_cur = ltp; ++nacc; for (int j = 0; j < P; ++j) { cur.th[j] = thp[j]; if (kind != 0) cur.g[j] = gp[j]; } }
      if (t >= p.n_burnin && (t - p.n_burnin) % thin == 0) {
        const int64_t s = (t - p.n_burnin) / thin;
        if (p.out_samples) for (int j = 0; j < P; ++j) ((T*)p.out_samples)[s * p.ss_iter + c * p.ss_chain + j * p.ss_param] = cur.th[j];
        if (p.out_target) ((T*)p.out_target)[s * p.n_chains + c] = lt_cur;
        if (p.out_accepted) p.out_accepted[s * p.n_chains + c] = acc ? 1 : 0;
      }
    }
    target[c] = lt_cur;
    if (p.accept_count) p.accept_count[c] += nacc;
  }
}

extern "C" {
int hostsim_gen_eval(int nl, const int* dims, const int* bias, const int* act, int loss, int dtype, int64_t C, const void* theta,
                     const void* x, const void* y, int64_t N, const void* loc, const void* scale, int has_t, double temp,
                     void* out_lt, void* out_g) {
  GenNet n = make_gen_net(nl, dims, bias, act, loss);
  if (dtype == EEYORE_B200_F64) gen_eval_all<double>(n, C, (const double*)theta, (const double*)x, (const double*)y, (int)N, (const double*)loc, (const double*)scale, has_t, temp, (double*)out_lt, (double*)out_g);
  else gen_eval_all<float>(n, C, (const float*)theta, (const float*)x, (const float*)y, (int)N, (const float*)loc, (const float*)scale, has_t, temp, (float*)out_lt, (float*)out_g);
  return n.P;
}
int hostsim_gen_run(int kind, int nl, const int* dims, const int* bias, const int* act, int loss, int dtype,
                    const eeyore_b200_run_params* p) {
  GenNet n = make_gen_net(nl, dims, bias, act, loss);
  if (dtype == EEYORE_B200_F64) gen_run_all<double>(n, kind, *p); else gen_run_all<float>(n, kind, *p);
  return 0;
}
}
