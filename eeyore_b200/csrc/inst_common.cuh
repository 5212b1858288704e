// Instantiation helper: builds the NetEntry of one architecture for fp32 and fp64.
#pragma once
#include "chain_kernels.cuh"
#include "smmala.cuh"
#include "adaptive.cuh"
#include "registry.h"

namespace eb {

template <typename T, class NET> cudaError_t eval_entry(const EvalCall& c) {
  ChainArgs<T> a{};
  a.n_chains = c.n_chains;
  a.theta = (T*)c.theta; a.x = (const T*)c.x; a.y = (const T*)c.y; a.n_rows = (int)c.n_rows;
  a.ploc = (const T*)c.ploc; a.pscale = (const T*)c.pscale;
  a.has_temperature = c.has_temperature; a.temperature = (T)c.temperature;
  a.target = (T*)c.out_target; a.grad = (T*)c.out_grad; a.out_ll = (T*)c.out_ll; a.out_lp = (T*)c.out_lp;
  a.use_bulk = c.use_bulk;
  return launch_eval<T, NET>(c.lanes, a, c.stream);
}

template <typename T> ChainArgs<T> chain_args_from(const eeyore_b200_run_params& p, int use_bulk, long NetP) {
  ChainArgs<T> a{};
  a.n_chains = p.n_chains; a.n_iters = p.n_iters; a.n_burnin = p.n_burnin; a.thin = p.thin < 1 ? 1 : p.thin;
  a.step = (T)p.step; a.num_steps = p.num_steps; a.symmetric = p.symmetric;
  a.has_temperature = p.has_temperature; a.temperature = (T)p.temperature;
  a.rng_mode = p.rng_mode;
  a.key = RngKey{(uint32_t)(p.seed & 0xffffffffu), (uint32_t)(p.seed >> 32)};
  a.iter0 = (uint32_t)p.iter_offset; a.chain0 = (uint32_t)p.chain_offset;
  a.z_tape = (const T*)p.z_tape; a.u_tape = (const T*)p.u_tape;
  a.x = (const T*)p.x; a.y = (const T*)p.y; a.n_rows = (int)p.n_rows;
  a.ploc = (const T*)p.prior_loc; a.pscale = (const T*)p.prior_scale;
  a.theta = (T*)p.theta; a.target = (T*)p.target; a.grad = (T*)p.grad;
  a.st_c = p.st_chain; a.st_p = p.st_param;
  if (a.st_c == 0 && a.st_p == 0) { a.st_c = NetP; a.st_p = 1; }
  a.out_samples = (T*)p.out_samples; a.ss_i = p.ss_iter; a.ss_c = p.ss_chain; a.ss_p = p.ss_param;
  a.out_target = (T*)p.out_target; a.out_grad = (T*)p.out_grad; a.out_acc = p.out_accepted;
  a.acc_count = p.accept_count;
  a.final_theta = (T*)p.final_theta; a.fs_c = p.fs_chain; a.fs_p = p.fs_param;
  a.final_target = (T*)p.final_target; a.final_acc = p.final_accept_count;
  a.use_bulk = use_bulk;
  a.tuner = DaTuner{p.tuner_l, p.tuner_d, p.tuner_m, p.tuner_logeub, p.tuner_has_eub};
  a.tuner_iter0 = p.tuner_iter0; a.tuner_burnin = p.tuner_burnin; a.tuner_state = p.tuner_state;
  return a;
}

template <typename T, class NET>
cudaError_t sampler_entry(int kind, const eeyore_b200_run_params& p, int lanes, int use_bulk) {
  return launch_sampler<T, NET>(kind, lanes, chain_args_from<T>(p, use_bulk, NET::P), (cudaStream_t)p.stream);
}

template <typename T, class NET>
cudaError_t forward_entry(int64_t n_chains, const void* theta, const void* x, int64_t n_rows, void* out,
                          cudaStream_t st) {
  ChainArgs<T> a{};
  a.n_chains = n_chains; a.theta = (T*)theta; a.x = (const T*)x; a.n_rows = (int)n_rows;
  return launch_forward<T, NET>(a, (T*)out, st);
}

template <typename T, class NET> cudaError_t smmala_entry(const eeyore_b200_run_params& p, int use_bulk) {
  if constexpr (NET::LOSS == LOSS_BINARY) {
    return launch_smmala<T, NET>(chain_args_from<T>(p, use_bulk, NET::P), (cudaStream_t)p.stream);
  } else {
    return cudaErrorNotSupported;
  }
}

template <typename T, class NET> cudaError_t adaptive_entry(int kind, const eeyore_b200_run_params& p, int use_bulk) {
  if constexpr (NET::P <= 32) {
    AdaptArgs ad{p.adapt_p[0], p.adapt_p[1], p.adapt_p[2], p.adapt_t0, (long)p.adapt_iter0, p.adapt_state, p.adapt_cov0,
                 p.adapt_status};
    return launch_adaptive<T, NET>(kind, chain_args_from<T>(p, use_bulk, NET::P), ad, (cudaStream_t)p.stream);
  } else {
    return cudaErrorNotSupported;
  }
}

template <typename T, class NET> NetEntry make_entry(int dtype) {
  NetEntry e{};
  e.n_layers = NET::NL;
  e.dims[0] = NET::D0; e.dims[1] = NET::D1; e.dims[2] = NET::D2; e.dims[3] = NET::D3;
  e.loss = NET::LOSS; e.dtype = dtype; e.n_params = NET::P;
  e.eval = &eval_entry<T, NET>;
  e.sampler = &sampler_entry<T, NET>;
  e.forward = &forward_entry<T, NET>;
  e.smmala = NET::LOSS == LOSS_BINARY ? &smmala_entry<T, NET> : nullptr;
  e.adaptive = NET::P <= 32 ? &adaptive_entry<T, NET> : nullptr;
  return e;
}

}  // namespace eb

#define EB_INSTANTIATE_NET(NAME, ...)                                                        \
  namespace eb {                                                                             \
  using Net_##NAME = Net<__VA_ARGS__>;                                                       \
  extern const NetEntry kNet_##NAME##_f32 = make_entry<float, Net_##NAME>(EEYORE_B200_F32);  \
  extern const NetEntry kNet_##NAME##_f64 = make_entry<double, Net_##NAME>(EEYORE_B200_F64); \
  }
