// extern "C" entry points of libeeyore_b200.so (declared in include/eeyore_b200.h).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include "registry.h"
#include "philox.cuh"
#include "generic.cuh"

namespace eb {
extern const NetEntry kNet_221_f32, kNet_221_f64, kNet_2321_f32, kNet_2321_f64, kNet_433_f32, kNet_433_f64,
    kNet_4323_f32, kNet_4323_f64;
static const NetEntry* const kNets[] = {&kNet_221_f32,  &kNet_221_f64, &kNet_2321_f32, &kNet_2321_f64,
                                        &kNet_433_f32,  &kNet_433_f64, &kNet_4323_f32, &kNet_4323_f64};

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
static int cuda_fail(cudaError_t e, const char* where) {
  return fail(EEYORE_B200_ECUDA, std::string(where) + ": " + cudaGetErrorString(e));
}
static const char* kTooLarge =
    "the data set does not fit the chain kernels' shared-memory staging (n_rows * (d0 + 1) values next to the kernel's "
    "tables; about 190 KB per CTA) or the network is too large for the runtime-shape kernels; the data-parallel path "
    "(eeyore_b200_dp_*) serves large data sets";
// Philox counters are 32-bit words (csrc/philox.cuh: block, iteration, chain id, kind): offsets beyond 2^32 would alias streams
static bool offsets_ok(const eeyore_b200_run_params* p) {
  const uint64_t lim = 1ull << 32;
  return p->chain_offset < lim && p->iter_offset < lim && p->chain_offset + (uint64_t)p->n_chains <= lim &&
         p->iter_offset + (uint64_t)p->n_iters <= lim;
}
static bool lanes_ok(int lanes) { return lanes == 0 || lanes == 1 || lanes == 4 || lanes == 8 || lanes == 16 || lanes == 32; }
static int use_bulk() {
  static int v = -1;
  if (v < 0) { const char* s = getenv("EEYORE_B200_NO_BULK"); v = (s && s[0] == '1') ? 0 : 1; }
  return v;
}
// lanes per chain: enough threads to fill the 148 SMs, never more lanes than data rows
static int choose_lanes(int64_t n_chains, int64_t n_rows, int requested) {
  if (requested > 0) return requested;
  // measured on B200 (config 2: 4096 chains, N = 150, P = 27, fp64): 8 lanes/chain 2.7e8 evals/s, 4 lanes 2.1e8,
  // 16 lanes 2.0e8, 32 lanes 1.4e8 -- the replicated per-lane work (prior, leapfrog, all-reduce) outweighs occupancy
  const int64_t want = 148LL * 192;
  const int cands[] = {1, 4, 8, 16, 32};
  int g = 1;
  for (int c : cands) {
    if (c > n_rows && c > 1) break;
    g = c;
    if (n_chains * c >= want) break;
  }
  return g;
}
}  // namespace eb

namespace eb {
// runtime-shape fallback (generic.cu)
cudaError_t generic_eval(const GenNet& net, int dtype, const EvalCall& c, void* out_fwd);
cudaError_t generic_sampler(const GenNet& net, int dtype, int kind, const eeyore_b200_run_params& p);
}  // namespace eb

struct eeyore_b200_mlp {
  const eb::NetEntry* net;   // compile-time specialisation, or nullptr -> runtime-shape path
  eb::GenNet gen;
  int dtype, n_params;
};

using namespace eb;

// forward pass of the runtime-shape path: the staging code reads a prior, so a unit prior is provided
static cudaError_t generic_forward_(eeyore_b200_mlp_t h, EvalCall c, void* out) {
  const size_t bytes = (size_t)h->gen.P * (h->dtype == EEYORE_B200_F64 ? 8 : 4);
  void* ones = nullptr;
  cudaError_t e = cudaMallocAsync(&ones, bytes, c.stream);
  if (e != cudaSuccess) return e;
  // any finite positive scale works: 0x3c.. patterns are avoided by filling with 1.0 through a tiny kernel-free trick:
  // cudaMemsetAsync cannot write 1.0, so fill on the host side
  std::string host(bytes, 0);
  if (h->dtype == EEYORE_B200_F64) { double* p = (double*)host.data(); for (int j = 0; j < h->gen.P; ++j) p[j] = 1.0; }
  else { float* p = (float*)host.data(); for (int j = 0; j < h->gen.P; ++j) p[j] = 1.0f; }
  e = cudaMemcpyAsync(ones, host.data(), bytes, cudaMemcpyHostToDevice, c.stream);
  if (e == cudaSuccess) { cudaStreamSynchronize(c.stream); c.ploc = ones; c.pscale = ones; c.y = c.x; e = generic_eval(h->gen, h->dtype, c, out); }
  cudaFreeAsync(ones, c.stream);
  return e;
}

extern "C" {

// shared with the other translation units of the library (not part of the public header)
int eeyore_b200_set_error_(int code, const char* msg) { return fail(code, msg); }

const char* eeyore_b200_last_error(void) { return g_err.c_str(); }
const char* eeyore_b200_version(void) { return "eeyore_b200 0.1 (sm_100a)"; }

int eeyore_b200_mlp_create(int n_layers, const int* dims, const int* bias, const int* act_ids, int loss_id, int dtype,
                           eeyore_b200_mlp_t* out) {
  if (!dims || !bias || !act_ids || !out) return fail(EEYORE_B200_EINVAL, "null argument");
  if (n_layers < 1) return fail(EEYORE_B200_EINVAL, "at least one dense layer is needed");  // 1 = LogisticRegression
  if (dtype != EEYORE_B200_F32 && dtype != EEYORE_B200_F64) return fail(EEYORE_B200_EINVAL, "dtype must be f32 or f64");
  if (loss_id != EEYORE_B200_LOSS_BINARY && loss_id != EEYORE_B200_LOSS_MULTICLASS)
    return fail(EEYORE_B200_EINVAL, "unknown loss id");
  if (n_layers > kGenMaxLayers) return fail(EEYORE_B200_EUNSUPPORTED, "at most 8 dense layers");
  for (int l = 0; l <= n_layers; ++l)
    if (dims[l] < 1) return fail(EEYORE_B200_EINVAL, "layer widths must be positive");
  for (int l = 0; l < n_layers; ++l)
    if (act_ids[l] != EEYORE_B200_ACT_NONE && act_ids[l] != EEYORE_B200_ACT_SIGMOID)
      return fail(EEYORE_B200_EUNSUPPORTED, "supported activations: torch.sigmoid and None");
  // the binary loss is evaluated on probabilities (sigmoid head, one output); the multiclass loss on logits (None head)
  if (loss_id == EEYORE_B200_LOSS_BINARY && (act_ids[n_layers - 1] != EEYORE_B200_ACT_SIGMOID || dims[n_layers] != 1))
    return fail(EEYORE_B200_EUNSUPPORTED, "binary_classification needs a single sigmoid output unit");
  if (loss_id == EEYORE_B200_LOSS_MULTICLASS && act_ids[n_layers - 1] != EEYORE_B200_ACT_NONE)
    return fail(EEYORE_B200_EUNSUPPORTED, "multiclass_classification needs a None (logit) head");
  bool standard = true;   // all biases on, sigmoid hidden units: eligible for a compile-time specialisation
  for (int l = 0; l < n_layers; ++l) standard = standard && bias[l] && (l == n_layers - 1 || act_ids[l] == EEYORE_B200_ACT_SIGMOID);
  auto* h = new eeyore_b200_mlp{};
  h->net = nullptr; h->dtype = dtype;
  h->gen = make_gen_net(n_layers, dims, bias, act_ids, loss_id);
  h->n_params = h->gen.P;
  if (standard && n_layers <= 3 && !getenv("EEYORE_B200_FORCE_GENERIC")) {
    for (const NetEntry* e : kNets) {
      if (e->n_layers != n_layers || e->loss != loss_id || e->dtype != dtype) continue;
      bool same = true;
      for (int l = 0; l <= n_layers; ++l) same = same && (e->dims[l] == dims[l]);
      if (same) { h->net = e; break; }
    }
  }
  *out = h;
  return EEYORE_B200_OK;
}

int eeyore_b200_mlp_destroy(eeyore_b200_mlp_t h) { delete h; return EEYORE_B200_OK; }
int eeyore_b200_mlp_num_params(eeyore_b200_mlp_t h) { return h ? h->n_params : EEYORE_B200_EINVAL; }
/* 1 if the handle is served by a compile-time specialisation, 0 for the runtime-shape kernels */
int eeyore_b200_mlp_is_specialised(eeyore_b200_mlp_t h) { return h && h->net ? 1 : 0; }

int eeyore_b200_log_target_grad(eeyore_b200_mlp_t h, int64_t n_chains, const void* theta, const void* x, const void* y,
                                int64_t n_rows, const void* prior_loc, const void* prior_scale, int has_temperature,
                                double temperature, void* out_target, void* out_grad, void* out_loglik,
                                void* out_logprior, int lanes_per_chain, void* stream) {
  if (!h || !theta || !x || !y || !prior_loc || !prior_scale) return fail(EEYORE_B200_EINVAL, "null argument");
  if (n_chains < 1 || n_rows < 1) return fail(EEYORE_B200_EINVAL, "n_chains and n_rows must be positive");
  if (!lanes_ok(lanes_per_chain)) return fail(EEYORE_B200_EINVAL, "lanes_per_chain must be 0 (auto), 1, 4, 8, 16 or 32");
  EvalCall c{n_chains, theta, x, y, n_rows, prior_loc, prior_scale, has_temperature, temperature,
             out_target, out_grad, out_loglik, out_logprior,
             choose_lanes(n_chains, n_rows, lanes_per_chain), use_bulk(), (cudaStream_t)stream};
  cudaError_t e = h->net ? h->net->eval(c) : generic_eval(h->gen, h->dtype, c, nullptr);
  if (e == cudaErrorInvalidConfiguration) return fail(EEYORE_B200_EUNSUPPORTED, kTooLarge);
  if (e != cudaSuccess) return cuda_fail(e, "log_target_grad");
  return EEYORE_B200_OK;
}

int eeyore_b200_forward(eeyore_b200_mlp_t h, int64_t n_chains, const void* theta, const void* x, int64_t n_rows,
                        void* out, void* stream) {
  if (!h || !theta || !x || !out) return fail(EEYORE_B200_EINVAL, "null argument");
  if (n_chains < 1 || n_rows < 1) return fail(EEYORE_B200_EINVAL, "n_chains and n_rows must be positive");
  cudaError_t e;
  if (h->net) {
    e = h->net->forward(n_chains, theta, x, n_rows, out, (cudaStream_t)stream);
  } else {
    EvalCall c{n_chains, theta, x, x /*unused*/, n_rows, nullptr, nullptr, 0, 0.0, nullptr, nullptr, nullptr, nullptr, 1, 0,
               (cudaStream_t)stream};
    e = generic_forward_(h, c, out);
  }
  if (e == cudaErrorInvalidConfiguration) return fail(EEYORE_B200_EUNSUPPORTED, kTooLarge);
  if (e != cudaSuccess) return cuda_fail(e, "forward");
  return EEYORE_B200_OK;
}

int64_t eeyore_b200_num_saved(int64_t n_iters, int64_t n_burnin, int64_t thin) {
  if (thin < 1) thin = 1;
  if (n_iters <= n_burnin) return 0;
  return (n_iters - n_burnin + thin - 1) / thin;
}

static int run_common(eeyore_b200_mlp_t h, const eeyore_b200_run_params* p, int kind, const char* name) {
  if (!h || !p) return fail(EEYORE_B200_EINVAL, "null argument");
  if (p->n_chains < 1 || p->n_rows < 1 || p->n_iters < 0) return fail(EEYORE_B200_EINVAL, "bad sizes");
  if (!p->theta || !p->target || !p->x || !p->y || !p->prior_loc || !p->prior_scale)
    return fail(EEYORE_B200_EINVAL, "null state / data pointer");
  if (kind != KIND_MH && !p->grad) return fail(EEYORE_B200_EINVAL, "grad state required");
  if (p->rng_mode == EEYORE_B200_RNG_TAPE && (!p->z_tape || !p->u_tape))
    return fail(EEYORE_B200_EINVAL, "tape mode needs z_tape and u_tape");
  if (kind == KIND_HMC && !p->tuner_state && p->num_steps < 1) return fail(EEYORE_B200_EINVAL, "num_steps must be >= 1");
  if (!(p->step > 0)) return fail(EEYORE_B200_EINVAL, "step must be positive");
  if (!lanes_ok(p->lanes_per_chain)) return fail(EEYORE_B200_EINVAL, "lanes_per_chain must be 0 (auto), 1, 4, 8, 16 or 32");
  if (p->rng_mode == EEYORE_B200_RNG_PHILOX && !offsets_ok(p))
    return fail(EEYORE_B200_EINVAL, "chain_offset + n_chains and iter_offset + n_iters must stay below 2^32 (32-bit Philox counter words)");
  if (p->n_iters == 0) return EEYORE_B200_OK;
  const int lanes = choose_lanes(p->n_chains, p->n_rows, p->lanes_per_chain);
  if (kind == KIND_HMC && p->tuner_state != nullptr) {
    if (!(p->tuner_l > 0)) return fail(EEYORE_B200_EINVAL, "tuner_l must be positive");
    kind = KIND_HMC_TUNED;
  }
  cudaError_t e;
  if (h->net) {
    e = h->net->sampler(kind, *p, lanes, use_bulk());
  } else {
    if (kind == KIND_HMC_TUNED) return fail(EEYORE_B200_EUNSUPPORTED, "HMCDATuner needs a compiled network specialisation");
    e = generic_sampler(h->gen, h->dtype, kind, *p);
  }
  if (e == cudaErrorInvalidConfiguration) return fail(EEYORE_B200_EUNSUPPORTED, kTooLarge);
  if (e != cudaSuccess) return cuda_fail(e, name);
  return EEYORE_B200_OK;
}

int eeyore_b200_mh_run(eeyore_b200_mlp_t h, const eeyore_b200_run_params* p) { return run_common(h, p, KIND_MH, "mh_run"); }
int eeyore_b200_mala_run(eeyore_b200_mlp_t h, const eeyore_b200_run_params* p) { return run_common(h, p, KIND_MALA, "mala_run"); }
int eeyore_b200_hmc_run(eeyore_b200_mlp_t h, const eeyore_b200_run_params* p) { return run_common(h, p, KIND_HMC, "hmc_run"); }

int eeyore_b200_smmala_run(eeyore_b200_mlp_t h, const eeyore_b200_run_params* p) {
  if (!h || !p) return fail(EEYORE_B200_EINVAL, "null argument");
  if (!h->net || !h->net->smmala)
    return fail(EEYORE_B200_EUNSUPPORTED, "SMMALA is built for the compiled binary-classification networks only");
  if (p->n_chains < 1 || p->n_rows < 1 || p->n_iters < 0) return fail(EEYORE_B200_EINVAL, "bad sizes");
  if (!p->theta || !p->target || !p->grad || !p->x || !p->y || !p->prior_loc || !p->prior_scale)
    return fail(EEYORE_B200_EINVAL, "null state / data pointer");
  if (p->rng_mode == EEYORE_B200_RNG_TAPE && (!p->z_tape || !p->u_tape))
    return fail(EEYORE_B200_EINVAL, "tape mode needs z_tape and u_tape");
  if (!(p->step > 0)) return fail(EEYORE_B200_EINVAL, "step must be positive");
  if (p->rng_mode == EEYORE_B200_RNG_PHILOX && !offsets_ok(p))
    return fail(EEYORE_B200_EINVAL, "chain_offset + n_chains and iter_offset + n_iters must stay below 2^32 (32-bit Philox counter words)");
  if (p->n_iters == 0) return EEYORE_B200_OK;
  cudaError_t e = h->net->smmala(*p, use_bulk());
  if (e == cudaErrorInvalidConfiguration) return fail(EEYORE_B200_EUNSUPPORTED, kTooLarge);
  if (e != cudaSuccess) return cuda_fail(e, "smmala_run");
  return EEYORE_B200_OK;
}

static int adaptive_common(eeyore_b200_mlp_t h, const eeyore_b200_run_params* p, int kind, const char* name) {
  if (!h || !p) return fail(EEYORE_B200_EINVAL, "null argument");
  if (!h->net || !h->net->adaptive)
    return fail(EEYORE_B200_EUNSUPPORTED, "AM / RAM are built for the compiled network specialisations with at most 32 parameters");
  if (p->n_chains < 1 || p->n_rows < 1 || p->n_iters < 0) return fail(EEYORE_B200_EINVAL, "bad sizes");
  if (!p->theta || !p->target || !p->x || !p->y || !p->prior_loc || !p->prior_scale || !p->adapt_state || !p->adapt_status)
    return fail(EEYORE_B200_EINVAL, "null state / data pointer");
  if (kind == 0 && !p->adapt_cov0) return fail(EEYORE_B200_EINVAL, "AM needs cov0");
  if (p->rng_mode == EEYORE_B200_RNG_TAPE && (!p->z_tape || !p->u_tape))
    return fail(EEYORE_B200_EINVAL, "tape mode needs z_tape and u_tape");
  if (p->rng_mode == EEYORE_B200_RNG_PHILOX && !offsets_ok(p))
    return fail(EEYORE_B200_EINVAL, "chain_offset + n_chains and iter_offset + n_iters must stay below 2^32 (32-bit Philox counter words)");
  if (p->n_iters == 0) return EEYORE_B200_OK;
  cudaError_t e = h->net->adaptive(kind, *p, use_bulk());
  if (e == cudaErrorInvalidConfiguration) return fail(EEYORE_B200_EUNSUPPORTED, kTooLarge);
  if (e != cudaSuccess) return cuda_fail(e, name);
  return EEYORE_B200_OK;
}

int eeyore_b200_am_run(eeyore_b200_mlp_t h, const eeyore_b200_run_params* p) { return adaptive_common(h, p, 0, "am_run"); }
int eeyore_b200_ram_run(eeyore_b200_mlp_t h, const eeyore_b200_run_params* p) { return adaptive_common(h, p, 1, "ram_run"); }

int64_t eeyore_b200_adapt_state_len(eeyore_b200_mlp_t h, int kind) {
  if (!h) return -1;
  const int64_t P = eeyore_b200_mlp_num_params(h);
  return kind == 0 ? P + 2 * P * P + 1 : P * P;
}

}  // extern "C"

// ---- Philox draw export and FMA-peak microbenchmark -------------------------------------------------------------
namespace eb {

template <typename T>
__global__ void philox_draws_kernel(long n_chains, int P, RngKey key, uint32_t iter, uint32_t chain0, T* z, T* u) {
  const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if constexpr (sizeof(T) == 8) {   // the fp64 Box-Muller log reads the shared-memory tables
    exp_table_init();
    __syncthreads();
  }
  if (c >= n_chains) return;
  const uint32_t gc = chain0 + (uint32_t)c;
  constexpr int PER = sizeof(T) == 8 ? 2 : 4;
  for (int j = 0; j < (P + PER - 1) / PER; ++j) {
    U4 w = philox4x32_10(U4{(uint32_t)j, iter, gc, 0u}, key.k0, key.k1);
    T v[4];
    if constexpr (sizeof(T) == 8) {
      box_muller<T>(Uni<double>::from(w.x, w.y), Uni<double>::from(w.z, w.w), &v[0], &v[1]);
    } else {
      box_muller<T>(Uni<float>::from(w.x), Uni<float>::from(w.y), &v[0], &v[1]);
      box_muller<T>(Uni<float>::from(w.z), Uni<float>::from(w.w), &v[2], &v[3]);
    }
    for (int k = 0; k < PER; ++k)
      if (j * PER + k < P) z[c * P + j * PER + k] = v[k];
  }
  if (u) u[c] = philox_uniform<T>(key, gc, iter);
}

// 8 independent FMA chains per thread; enough warps to saturate every SMSP.
template <typename T> __global__ void fma_peak_kernel(int iters, T seed, T* sink) {
  T a0 = seed, a1 = seed + T(1), a2 = seed + T(2), a3 = seed + T(3), a4 = seed + T(4), a5 = seed + T(5),
    a6 = seed + T(6), a7 = seed + T(7);
  const T m = T(0.999999), c = T(1e-7);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      a0 = fma_t<T>(a0, m, c); a1 = fma_t<T>(a1, m, c); a2 = fma_t<T>(a2, m, c); a3 = fma_t<T>(a3, m, c);
      a4 = fma_t<T>(a4, m, c); a5 = fma_t<T>(a5, m, c); a6 = fma_t<T>(a6, m, c); a7 = fma_t<T>(a7, m, c);
    }
  }
  T s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == T(-12345)) sink[0] = s;
}

template <typename T> int fma_peak(int iters, double* out) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  T* sink = nullptr;
  cudaError_t e = cudaMalloc(&sink, sizeof(T));
  if (e != cudaSuccess) return cuda_fail(e, "fma_peak");
  const int blocks = sms * 8, threads = 256;
  cudaEvent_t t0, t1;
  cudaEventCreate(&t0); cudaEventCreate(&t1);
  fma_peak_kernel<T><<<blocks, threads>>>(iters / 4 + 1, T(1), sink);  // warm-up
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(t0);
    fma_peak_kernel<T><<<blocks, threads>>>(iters, T(1), sink);
    cudaEventRecord(t1);
    e = cudaEventSynchronize(t1);
    if (e != cudaSuccess) break;
    float ms = 0;
    cudaEventElapsedTime(&ms, t0, t1);
    const double flops = 2.0 * 8 * 16 * (double)iters * blocks * threads;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(t0); cudaEventDestroy(t1);
  cudaFree(sink);
  if (e != cudaSuccess) return cuda_fail(e, "fma_peak");
  *out = best;
  return EEYORE_B200_OK;
}
}  // namespace eb

extern "C" {

int eeyore_b200_philox_draws(int dtype, int64_t n_chains, int n_params, uint64_t seed, uint64_t iter,
                             uint64_t chain_offset, void* out_z, void* out_u, void* stream) {
  if (!out_z || n_chains < 1 || n_params < 1) return fail(EEYORE_B200_EINVAL, "bad argument");
  RngKey key{(uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32)};
  const int threads = 128;
  const unsigned blocks = (unsigned)((n_chains + threads - 1) / threads);
  if (dtype == EEYORE_B200_F64)
    philox_draws_kernel<double><<<blocks, threads, 0, (cudaStream_t)stream>>>(n_chains, n_params, key, (uint32_t)iter,
                                                                             (uint32_t)chain_offset, (double*)out_z, (double*)out_u);
  else if (dtype == EEYORE_B200_F32)
    philox_draws_kernel<float><<<blocks, threads, 0, (cudaStream_t)stream>>>(n_chains, n_params, key, (uint32_t)iter,
                                                                            (uint32_t)chain_offset, (float*)out_z, (float*)out_u);
  else return fail(EEYORE_B200_EINVAL, "dtype must be f32 or f64");
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "philox_draws");
  return EEYORE_B200_OK;
}

int eeyore_b200_fma_peak(int dtype, int iters, double* out_tflops) {
  if (!out_tflops || iters < 1) return fail(EEYORE_B200_EINVAL, "bad argument");
  if (dtype == EEYORE_B200_F64) return fma_peak<double>(iters, out_tflops);
  if (dtype == EEYORE_B200_F32) return fma_peak<float>(iters, out_tflops);
  return fail(EEYORE_B200_EINVAL, "dtype must be f32 or f64");
}

}  // extern "C"
