// Kernels of the runtime-shape path (generic.cuh): one thread per chain, per-thread vectors in shared memory.
#include <cuda_runtime.h>
#include "generic.cuh"
#include "registry.h"

namespace eb {

template <typename T> struct GenArgs {
  GenNet net;
  long n_chains, n_iters, n_burnin, thin;
  T step;
  int num_steps, symmetric, has_temperature, rng_mode;
  T temperature;
  RngKey key;
  uint32_t iter0, chain0;
  const T *z_tape, *u_tape, *x, *y;
  int n_rows;
  const T *ploc, *pscale;
  T *theta, *target, *grad;
  long st_c, st_p;
  T* out_samples;
  long ss_i, ss_c, ss_p;
  T *out_target, *out_grad;
  uint8_t* out_acc;
  uint32_t* acc_count;
  T *out_ll, *out_lp, *out_fwd;
};

__host__ __device__ inline size_t gen_al16(size_t v) { return (v + 15) & ~size_t(15); }

// shared memory: [x | y or cls | ploc | pivar | misc | per-thread vectors (n_vec * stride elements)]
template <typename T> struct GenLayout {
  size_t off_x, off_y, off_ploc, off_pivar, off_misc, off_vec, total;
  __host__ __device__ GenLayout(const GenNet& n, int n_rows, int threads, int n_vec_elems) {
    size_t o = 0;
    off_x = o; o += gen_al16(sizeof(T) * (size_t)n_rows * n.dims[0]);
    off_y = o; o += gen_al16((sizeof(T) > 4 ? sizeof(T) : 4) * (size_t)n_rows);
    off_ploc = o; o += gen_al16(sizeof(T) * n.P);
    off_pivar = o; o += gen_al16(sizeof(T) * n.P);
    off_misc = o; o += 16;
    off_vec = o; o += sizeof(T) * (size_t)n_vec_elems * threads;
    total = o;
  }
};

template <typename T> __device__ DataView<T> gen_stage(unsigned char* smem, const GenLayout<T>& lay, const GenArgs<T>& a) {
  const GenNet& n = a.net;
  T* xs = reinterpret_cast<T*>(smem + lay.off_x);
  T* ys = reinterpret_cast<T*>(smem + lay.off_y);
  int* cs = reinterpret_cast<int*>(smem + lay.off_y);
  T* ploc = reinterpret_cast<T*>(smem + lay.off_ploc);
  T* pivar = reinterpret_cast<T*>(smem + lay.off_pivar);
  T* misc = reinterpret_cast<T*>(smem + lay.off_misc);
  const int tid = threadIdx.x, N = a.n_rows, dl = n.dims[n.nl];
  if constexpr (sizeof(T) == 8) exp_table_init();
  for (int i = tid; i < N * n.dims[0]; i += blockDim.x) xs[i] = a.x[i];
  if (n.loss == LOSS_BINARY) {
    for (int i = tid; i < N; i += blockDim.x) ys[i] = a.y[i];
  } else {
    for (int i = tid; i < N; i += blockDim.x) {
      const T* row = a.y + (size_t)i * dl;
      int best = 0;
      T bv = row[0];
      for (int k = 1; k < dl; ++k) if (row[k] > bv) { bv = row[k]; best = k; }
      cs[i] = best;
    }
  }
  for (int j = tid; j < n.P; j += blockDim.x) {
    const T s = a.pscale[j];
    ploc[j] = a.ploc[j];
    pivar[j] = T(1) / (s * s);
  }
  if (tid == 0) {
    T c = T(0);
    for (int j = 0; j < n.P; ++j) c += -log_t<T>(a.pscale[j]) - T(kLogSqrt2Pi);
    misc[0] = c;
  }
  __syncthreads();
  DataView<T> d;
  d.hard_labels = false;   // the runtime-shape row code keeps its soft-label branch
  d.x = xs; d.y = ys; d.cls = cs; d.n_rows = N; d.ploc = ploc; d.pivar = pivar; d.lp_const = misc[0];
  d.temperature = a.temperature; d.has_temperature = a.has_temperature != 0;
  return d;
}

template <typename T> __device__ GenWork<T> gen_work(T* base, int stride, const GenNet& n, T** next) {
  GenWork<T> w;
  w.h = StridedVec<T>{base, stride};
  w.da = StridedVec<T>{base + (size_t)n.H * stride, stride};
  w.db = StridedVec<T>{base + (size_t)(n.H + n.maxd) * stride, stride};
  *next = base + (size_t)(n.H + 2 * n.maxd) * stride;
  return w;
}

// mode 0: log_target (+grad) ; mode 1: forward outputs [C, N, dL]
template <typename T> __global__ void gen_eval_kernel(const GenArgs<T> a, int mode) {
  extern __shared__ __align__(16) unsigned char smem[];
  const GenNet& n = a.net;
  const int stride = blockDim.x;
  const GenLayout<T> lay(n, a.n_rows, stride, n.H + 2 * n.maxd + 2 * n.P);
  const DataView<T> d = gen_stage<T>(smem, lay, a);
  long chain = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = chain < a.n_chains;
  if (!live) chain = a.n_chains - 1;
  T* vec = reinterpret_cast<T*>(smem + lay.off_vec) + threadIdx.x;
  T* nxt;
  const GenWork<T> w = gen_work<T>(vec, stride, n, &nxt);
  StridedVec<T> th{nxt, stride}, g{nxt + (size_t)n.P * stride, stride};
  for (int j = 0; j < n.P; ++j) th[j] = a.theta[chain * n.P + j];
  if (mode == 1) {
    const int dl = n.dims[n.nl];
    for (int i = 0; i < a.n_rows; ++i) {
      T ll = T(0);
      int dummy = 0;
      gen_accumulate_row<T, false>(n, th, d.x + (size_t)i * n.dims[0], T(0), 0, w, ll, dummy);
      if (live)
        for (int k = 0; k < dl; ++k) {
          const T v = w.h[n.hoff[n.nl] + k];
          a.out_fwd[(chain * a.n_rows + i) * dl + k] = n.act[n.nl - 1] ? sigmoid_t<T>(v) : v;
        }
    }
    return;
  }
  T lt, ll, lp;
  if (a.grad) gen_eval_target<T, true>(n, d, th, w, lt, g, &ll, &lp);
  else { int dummy = 0; gen_eval_target<T, false>(n, d, th, w, lt, dummy, &ll, &lp); }
  if (live) {
    if (a.target) a.target[chain] = lt;
    if (a.out_ll) a.out_ll[chain] = ll;
    if (a.out_lp) a.out_lp[chain] = lp;
    if (a.grad) for (int j = 0; j < n.P; ++j) a.grad[chain * n.P + j] = g[j];
  }
}

template <typename T, int KIND> __global__ void gen_sampler_kernel(const GenArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const GenNet& n = a.net;
  const int stride = blockDim.x, P = n.P;
  const GenLayout<T> lay(n, a.n_rows, stride, n.H + 2 * n.maxd + 3 * P);
  const DataView<T> d = gen_stage<T>(smem, lay, a);
  long chain = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = chain < a.n_chains;
  if (!live) chain = a.n_chains - 1;
  T* vec = reinterpret_cast<T*>(smem + lay.off_vec) + threadIdx.x;
  T* nxt;
  const GenWork<T> w = gen_work<T>(vec, stride, n, &nxt);
  StridedVec<T> z{nxt, stride}, thp{nxt + (size_t)P * stride, stride}, gp{nxt + (size_t)2 * P * stride, stride};
  Cur<T> cur;
  cur.th = a.theta + chain * a.st_c;
  cur.g = (KIND != KIND_MH) ? a.grad + chain * a.st_c : nullptr;
  cur.stride = a.st_p;
  T lt_cur = a.target[chain];
  const T step = a.step, half_step = T(0.5) * step, sd = sqrt_t<T>(step);
  const uint32_t gchain = a.chain0 + (uint32_t)chain;
  uint32_t n_acc = 0;
  for (long t = 0; t < a.n_iters; ++t) {
    T u, ltp;
    if (a.rng_mode == 0) {
      gen_philox_normals<T>(z, P, a.key, gchain, a.iter0 + (uint32_t)t);
      u = philox_uniform<T>(a.key, gchain, a.iter0 + (uint32_t)t);
    } else {
      const T* zt = a.z_tape + ((size_t)t * a.n_chains + chain) * P;
      for (int j = 0; j < P; ++j) z[j] = zt[j];
      u = a.u_tape[(size_t)t * a.n_chains + chain];
    }
    bool acc;
    if (KIND == KIND_MH) acc = gen_mh_draw<T>(n, d, w, step, a.symmetric != 0, cur, lt_cur, z, u, thp, ltp);
    else if (KIND == KIND_MALA) acc = gen_mala_draw<T>(n, d, w, half_step, sd, cur, lt_cur, z, u, thp, gp, ltp);
    else acc = gen_hmc_draw<T>(n, d, w, step, half_step, a.num_steps, cur, lt_cur, z, u, thp, gp, ltp);
    if (acc) {
      lt_cur = ltp;
      ++n_acc;
      if (live)
        for (int j = 0; j < P; ++j) {
          cur.th[j * cur.stride] = thp[j];
          if (KIND != KIND_MH) cur.g[j * cur.stride] = gp[j];
        }
    }
    if (t >= a.n_burnin && (t - a.n_burnin) % a.thin == 0 && live) {
      const long s = (t - a.n_burnin) / a.thin;
      if (a.out_samples) for (int j = 0; j < P; ++j) a.out_samples[s * a.ss_i + chain * a.ss_c + j * a.ss_p] = cur.th[j * cur.stride];
      if (KIND != KIND_MH && a.out_grad)
        for (int j = 0; j < P; ++j) a.out_grad[s * a.ss_i + chain * a.ss_c + j * a.ss_p] = cur.g[j * cur.stride];
      if (a.out_target) a.out_target[s * a.n_chains + chain] = lt_cur;
      if (a.out_acc) a.out_acc[s * a.n_chains + chain] = acc ? 1 : 0;
    }
  }
  if (live) {
    a.target[chain] = lt_cur;
    if (a.acc_count) a.acc_count[chain] += n_acc;
  }
}

// pick the widest CTA (128 ... 8 threads) whose per-thread vectors fit shared memory; 0 if even 8 threads do not fit
// (narrow CTAs are slow, but this path is the functional fallback for shapes without a compiled specialisation)
template <typename T> int gen_pick_threads(const GenNet& n, int n_rows, int n_vec_elems, size_t* smem_out) {
  int dev = 0, max_smem = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (sizeof(T) == 8) max_smem -= (int)(sizeof(double) * (kExpTabN + 2 * kLogTabN));   // the static fp64 math tables
  for (int th = 128; th >= 8; th >>= 1) {
    const GenLayout<T> lay(n, n_rows, th, n_vec_elems);
    if (lay.total <= (size_t)max_smem) { *smem_out = lay.total; return th; }
  }
  return 0;
}

template <typename T>
cudaError_t gen_launch_eval(const GenNet& net, const EvalCall& c, void* out_fwd) {
  GenArgs<T> a{};
  a.net = net; a.n_chains = c.n_chains;
  a.theta = (T*)c.theta; a.x = (const T*)c.x; a.y = (const T*)c.y; a.n_rows = (int)c.n_rows;
  a.ploc = (const T*)c.ploc; a.pscale = (const T*)c.pscale;
  a.has_temperature = c.has_temperature; a.temperature = (T)c.temperature;
  a.target = (T*)c.out_target; a.grad = (T*)c.out_grad; a.out_ll = (T*)c.out_ll; a.out_lp = (T*)c.out_lp;
  a.out_fwd = (T*)out_fwd;
  size_t smem = 0;
  const int threads = gen_pick_threads<T>(net, a.n_rows, net.H + 2 * net.maxd + 2 * net.P, &smem);
  if (!threads) return cudaErrorInvalidConfiguration;
  cudaError_t e = cudaFuncSetAttribute(gen_eval_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const long blocks = (a.n_chains + threads - 1) / threads;
  gen_eval_kernel<T><<<(unsigned)blocks, threads, smem, c.stream>>>(a, out_fwd ? 1 : 0);
  return cudaGetLastError();
}

template <typename T> cudaError_t gen_launch_sampler(const GenNet& net, int kind, const eeyore_b200_run_params& p) {
  GenArgs<T> a{};
  a.net = net;
  a.n_chains = p.n_chains; a.n_iters = p.n_iters; a.n_burnin = p.n_burnin; a.thin = p.thin < 1 ? 1 : p.thin;
  a.step = (T)p.step; a.num_steps = p.num_steps; a.symmetric = p.symmetric;
  a.has_temperature = p.has_temperature; a.temperature = (T)p.temperature; a.rng_mode = p.rng_mode;
  a.key = RngKey{(uint32_t)(p.seed & 0xffffffffu), (uint32_t)(p.seed >> 32)};
  a.iter0 = (uint32_t)p.iter_offset; a.chain0 = (uint32_t)p.chain_offset;
  a.z_tape = (const T*)p.z_tape; a.u_tape = (const T*)p.u_tape;
  a.x = (const T*)p.x; a.y = (const T*)p.y; a.n_rows = (int)p.n_rows;
  a.ploc = (const T*)p.prior_loc; a.pscale = (const T*)p.prior_scale;
  a.theta = (T*)p.theta; a.target = (T*)p.target; a.grad = (T*)p.grad;
  a.st_c = p.st_chain; a.st_p = p.st_param;
  if (a.st_c == 0 && a.st_p == 0) { a.st_c = net.P; a.st_p = 1; }
  a.out_samples = (T*)p.out_samples; a.ss_i = p.ss_iter; a.ss_c = p.ss_chain; a.ss_p = p.ss_param;
  a.out_target = (T*)p.out_target; a.out_grad = (T*)p.out_grad; a.out_acc = p.out_accepted; a.acc_count = p.accept_count;
  size_t smem = 0;
  const int threads = gen_pick_threads<T>(net, a.n_rows, net.H + 2 * net.maxd + 3 * net.P, &smem);
  if (!threads) return cudaErrorInvalidConfiguration;
  const long blocks = (a.n_chains + threads - 1) / threads;
  cudaStream_t st = (cudaStream_t)p.stream;
  cudaError_t e;
#define EB_GEN_LAUNCH(K)                                                                                   \
  e = cudaFuncSetAttribute(gen_sampler_kernel<T, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
  if (e != cudaSuccess) return e;                                                                          \
  gen_sampler_kernel<T, K><<<(unsigned)blocks, threads, smem, st>>>(a);
  if (kind == KIND_MH) { EB_GEN_LAUNCH(KIND_MH) }
  else if (kind == KIND_MALA) { EB_GEN_LAUNCH(KIND_MALA) }
  else { EB_GEN_LAUNCH(KIND_HMC) }
#undef EB_GEN_LAUNCH
  return cudaGetLastError();
}

// type-erased entry points used by capi.cu
cudaError_t generic_eval(const GenNet& net, int dtype, const EvalCall& c, void* out_fwd) {
  return dtype == EEYORE_B200_F64 ? gen_launch_eval<double>(net, c, out_fwd) : gen_launch_eval<float>(net, c, out_fwd);
}
cudaError_t generic_sampler(const GenNet& net, int dtype, int kind, const eeyore_b200_run_params& p) {
  return dtype == EEYORE_B200_F64 ? gen_launch_sampler<double>(net, kind, p) : gen_launch_sampler<float>(net, kind, p);
}

}  // namespace eb
