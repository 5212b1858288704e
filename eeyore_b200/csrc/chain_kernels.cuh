// Chain-batched kernels: thread mapping, shared-memory staging (cp.async.bulk / TMA 1-D bulk copy), on-device
// Philox or tape noise, fused sampler loops with sample write-out.  One launch runs all iterations of all chains:
// an entire MCMC run (every HMC trajectory of it) executes without leaving the SM.
//
// Replaces the loop of eeyore/samplers/serial_sampler.py:35-52 around <Sampler>.draw, for C independent chains
// (the semantics of SerialSampler.benchmark, serial_sampler.py:54-126), and the model evaluation
// eeyore/models/log_target_model.py:20-23.
#pragma once
#include "samplers.cuh"
#include "philox.cuh"
#include "registry.h"
#include <type_traits>

namespace eb {

constexpr int kBlock = 128;

template <typename T> struct ChainArgs {
  long n_chains, n_iters, n_burnin, thin;
  T step;
  int num_steps, symmetric, has_temperature, rng_mode;
  T temperature;
  RngKey key;
  uint32_t iter0, chain0;
  const T* z_tape;
  const T* u_tape;
  const T* x;
  const T* y;
  int n_rows;
  const T* ploc;
  const T* pscale;
  T* theta;
  T* target;
  T* grad;
  long st_c, st_p;  // state layout: element (c, j) at c*st_c + j*st_p
  T* out_samples;
  long ss_i, ss_c, ss_p;
  T* out_target;
  T* out_grad;
  uint8_t* out_acc;
  uint32_t* acc_count;
  DaTuner tuner;       // HMC only; tuner_state == nullptr disables
  long tuner_iter0, tuner_burnin;
  double* tuner_state;
  T* out_ll;   // eval kernel only
  T* out_lp;   // eval kernel only
  int use_bulk;
  T* final_theta;      // sampler kernels: final state once more (host-visible copy), or nullptr
  long fs_c, fs_p;
  T* final_target;
  uint32_t* final_acc;
};

__host__ __device__ inline size_t align16(size_t v) { return (v + 15) & ~size_t(15); }

// shared-memory layout (bytes), identical on host (size) and device (carve)
template <typename T, class NET> struct SmemLayout {
  size_t off_bar, off_x, off_y, off_ploc, off_pivar, off_misc, off_mom, off_grad, total;
  __host__ __device__ SmemLayout(int n_rows, int chains_per_block, bool with_momentum, bool with_grad = false) {
    size_t o = 0;
    off_bar = o; o += 16;
    off_x = o; o += align16(sizeof(T) * (size_t)n_rows * NET::D0);
    off_y = o; o += align16((sizeof(T) > 4 ? sizeof(T) : 4) * (size_t)n_rows);
    off_ploc = o; o += align16(sizeof(T) * NET::P);
    off_pivar = o; o += align16(sizeof(T) * NET::P);
    off_misc = o; o += 16;
    off_mom = o; if (with_momentum) o += align16(sizeof(T) * NET::P * (size_t)chains_per_block);
    off_grad = o; if (with_grad) o += align16(sizeof(T) * NET::P * (size_t)kBlock);
    total = o;
  }
};

// Resident blocks per SM the sampler kernels are compiled for (register budget = 65536 / (kBlock * blocks)):
// theta' and the gradient accumulators (2 P values) must stay in registers.
// fp64 gradient accumulators of the gradient-based samplers live in shared memory (one column per thread) when the
// parameter vector is large enough that theta' + accumulators would not fit a 128-register budget.
template <typename T, class NET, int KIND> constexpr bool grad_in_smem() {
  // measured on B200 (config 4): 128 registers + shared-memory accumulators = 9.3e9 evals/s, 168 registers with
  // register accumulators = 10.4e9 evals/s  -> kept off by default
#ifdef EB_GRAD_IN_SMEM
  return sizeof(T) == 8 && NET::P >= 16 && KIND != KIND_MH;
#else
  return false;
#endif
}
// two resident CTAs (255 registers) since the fp64 fast path interleaves two rows per lane (mlp_static.cuh: EB_ROW_BATCH)
#ifndef EB_MINB_F64_SMALL
#define EB_MINB_F64_SMALL 2
#endif
#ifndef EB_MINB_F64_LARGE
#define EB_MINB_F64_LARGE 2
#endif
// Resident CTAs per SM the sampler kernels are compiled for (register budget = 65536 / (kBlock * blocks)).
template <typename T, class NET, int KIND> constexpr int min_blocks() {
  if (sizeof(T) == 4) return NET::P <= 32 ? 4 : 3;
  if (grad_in_smem<T, NET, KIND>()) return NET::P <= 20 ? 4 : 3;
  return NET::P <= 20 ? EB_MINB_F64_SMALL : (NET::P <= 32 ? EB_MINB_F64_LARGE : 2);
}

// Threads per CTA of the sampler kernels.  fp64 / small P runs two CTAs per SM for the register budget of the two-row fast
// path; EB_SAMPLER_BLOCK_F64_SMALL > 128 puts more warps into those two CTAs (budget = 65536 / (2 * threads)).
#ifndef EB_SAMPLER_BLOCK_F64_SMALL
#define EB_SAMPLER_BLOCK_F64_SMALL 128
#endif
template <typename T, class NET, int KIND> constexpr int sampler_block() {
  return (sizeof(T) == 8 && NET::P <= 20 && !grad_in_smem<T, NET, KIND>()) ? EB_SAMPLER_BLOCK_F64_SMALL : kBlock;
}

// ---- TMA 1-D bulk copy global -> shared, completion on an mbarrier --------------------------------------------
EB_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

EB_D void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
EB_D void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
EB_D void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
EB_D void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra WAIT_%=;\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// Stages x, y (or class labels), prior into shared memory and returns the block's DataView.
template <typename T, class NET>
EB_D DataView<T> stage_data(unsigned char* smem, const SmemLayout<T, NET>& lay, const ChainArgs<T>& a) {
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + lay.off_bar);
  T* xs = reinterpret_cast<T*>(smem + lay.off_x);
  T* ys = reinterpret_cast<T*>(smem + lay.off_y);
  int* cs = reinterpret_cast<int*>(smem + lay.off_y);
  T* ploc = reinterpret_cast<T*>(smem + lay.off_ploc);
  T* pivar = reinterpret_cast<T*>(smem + lay.off_pivar);
  T* misc = reinterpret_cast<T*>(smem + lay.off_misc);
  const int tid = threadIdx.x;
  const int N = a.n_rows;
  if constexpr (sizeof(T) == 8) exp_table_init();  // visible after the __syncthreads() below

  const uint32_t x_bytes = (uint32_t)(sizeof(T) * (size_t)N * NET::D0);
  const uint32_t y_bytes = (uint32_t)(sizeof(T) * (size_t)N);
  uint32_t xb = 0, yb = 0;  // bytes moved by the bulk engine
  if (a.use_bulk) {
    if ((reinterpret_cast<uintptr_t>(a.x) & 15) == 0) xb = x_bytes & ~15u;
    if (NET::LOSS == LOSS_BINARY && (reinterpret_cast<uintptr_t>(a.y) & 15) == 0) yb = y_bytes & ~15u;
  }
  if (xb + yb > 0) {
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(bar, xb + yb);
      if (xb) bulk_g2s(xs, a.x, xb, bar);
      if (yb) bulk_g2s(ys, a.y, yb, bar);
    }
  }
  // tails (and everything, when the bulk path is off or the source is unaligned)
  for (int i = xb / sizeof(T) + tid; i < N * NET::D0; i += blockDim.x) xs[i] = a.x[i];
  if constexpr (NET::LOSS == LOSS_BINARY) {
    for (int i = yb / sizeof(T) + tid; i < N; i += blockDim.x) ys[i] = a.y[i];
  } else {
    // torch.argmax(y, 1) of the one-hot row (first maximal index), constants.py:17
    for (int i = tid; i < N; i += blockDim.x) {
      const T* row = a.y + (size_t)i * NET::DL;
      int best = 0;
      T bv = row[0];
      for (int k = 1; k < NET::DL; ++k) if (row[k] > bv) { bv = row[k]; best = k; }
      cs[i] = best;
    }
  }
  for (int j = tid; j < NET::P; j += blockDim.x) {
    const T s = a.pscale[j];
    ploc[j] = a.ploc[j];
    pivar[j] = T(1) / (s * s);
  }
  if (tid == 0) {
    T c = T(0);
    for (int j = 0; j < NET::P; ++j) c += -log_t<T>(a.pscale[j]) - T(kLogSqrt2Pi);
    misc[0] = c;
  }
  if (xb + yb > 0) mbar_wait(bar, 0);
  __syncthreads();
  int hard = 1;
  if constexpr (NET::LOSS == LOSS_BINARY) {
    for (int i = tid; i < N; i += blockDim.x) hard &= (ys[i] == T(0) || ys[i] == T(1)) ? 1 : 0;
    hard = __syncthreads_and(hard);
  }
  DataView<T> d;
  d.hard_labels = hard != 0;
  d.x = xs; d.y = ys; d.cls = cs; d.n_rows = N; d.ploc = ploc; d.pivar = pivar; d.lp_const = misc[0];
  d.temperature = a.temperature; d.has_temperature = a.has_temperature != 0;
  return d;
}

// ---- log_target / gradient of C chains (one evaluation) ---------------------------------------------------------
template <typename T, class NET, int G>
__global__ void __launch_bounds__(kBlock) eval_kernel(const ChainArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int CPB = kBlock / G;
  const SmemLayout<T, NET> lay(a.n_rows, CPB, false);
  const DataView<T> d = stage_data<T, NET>(smem, lay, a);
  const int sub = threadIdx.x % G;
  long chain = (long)blockIdx.x * CPB + threadIdx.x / G;
  const bool live = chain < a.n_chains;
  if (!live) chain = a.n_chains - 1;  // keep every lane in the shuffles
  T th[NET::P], g[NET::P];
#pragma unroll
  for (int j = 0; j < NET::P; ++j) th[j] = a.theta[chain * NET::P + j];
  T lt, ll, lp;
  if (a.grad != nullptr) eval_target<T, NET, G, true>(d, sub, th, lt, g, &ll, &lp);
  else { int dummy = 0; eval_target<T, NET, G, false>(d, sub, th, lt, dummy, &ll, &lp); }
  if (live && sub == 0) {
    if (a.target) a.target[chain] = lt;
    if (a.out_ll) a.out_ll[chain] = ll;
    if (a.out_lp) a.out_lp[chain] = lp;
    if (a.grad) {
#pragma unroll
      for (int j = 0; j < NET::P; ++j) a.grad[chain * NET::P + j] = g[j];
    }
  }
}

// ---- MLP.forward for C chains: out [C, N, DL] ---------------------------------------------------------------------
template <typename T, class NET>
__global__ void __launch_bounds__(kBlock) forward_kernel(const ChainArgs<T> a, T* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem[];
  const long chain = (long)blockIdx.x * kBlock + threadIdx.x;
  T* xs = reinterpret_cast<T*>(smem);
  if constexpr (sizeof(T) == 8) exp_table_init();
  for (int i = threadIdx.x; i < a.n_rows * NET::D0; i += blockDim.x) xs[i] = a.x[i];
  __syncthreads();
  if (chain >= a.n_chains) return;
  T th[NET::P];
#pragma unroll
  for (int j = 0; j < NET::P; ++j) th[j] = a.theta[chain * NET::P + j];
  for (int i = 0; i < a.n_rows; ++i) {
    T o[NET::DL];
    forward_row<T, NET>(th, xs + i * NET::D0, o);
#pragma unroll
    for (int k = 0; k < NET::DL; ++k) out[(chain * a.n_rows + i) * NET::DL + k] = o[k];
  }
}

// ---- fused sampler: all iterations of all chains in one launch --------------------------------------------------
// Register-resident per lane: the proposal theta' and the gradient accumulators.  Shared memory: data set, prior and
// (HMC) the momentum.  Global memory (coalesced in the chain-minor layout): the chain's current sample / gradient, read
// once per iteration and written on accept.
template <typename T, class NET, int G, int KIND>
#ifdef EB_MAXNREG   // experiment: an explicit register cap instead of the launch bounds.  Config 4 on B200 (1.753e10 evals/s with the
                    // launch bounds, 236 registers): cap 224 -> 1.720e10, 192 -> 1.556e10, 168 (3 CTAs per SM, 116 B of spills) -> 1.576e10
__global__ void __maxnreg__(EB_MAXNREG) sampler_kernel(const ChainArgs<T> a) {
#else
__global__ void __launch_bounds__(sampler_block<T, NET, KIND>(), min_blocks<T, NET, KIND>()) sampler_kernel(const ChainArgs<T> a) {
#endif
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int KB = sampler_block<T, NET, KIND>();
  constexpr int CPB = KB / G;
  constexpr int P = NET::P;
  constexpr bool IS_HMC = KIND == KIND_HMC || KIND == KIND_HMC_TUNED;
  constexpr bool GSM = grad_in_smem<T, NET, KIND>();
#ifdef EB_THETA_IN_SMEM
  const SmemLayout<T, NET> lay(a.n_rows, CPB, IS_HMC, true);
#else
  const SmemLayout<T, NET> lay(a.n_rows, CPB, IS_HMC, GSM);
#endif
  const DataView<T> d = stage_data<T, NET>(smem, lay, a);
  const int sub = threadIdx.x % G;
  const int cl = threadIdx.x / G;
  long chain = (long)blockIdx.x * CPB + cl;
  const bool live = chain < a.n_chains;
  // lanes of a padding group shadow the last chain (so that every lane takes part in the shuffles) but never write
  if (!live) chain = a.n_chains - 1;

  Cur<T> cur;
  cur.th = a.theta + chain * a.st_c;
  cur.g = (KIND != KIND_MH) ? a.grad + chain * a.st_c : nullptr;
  cur.stride = a.st_p;
  T* mom = reinterpret_cast<T*>(smem + lay.off_mom) + cl;
  T lt_cur = a.target[chain];

  T step = a.step;
  T half_step = T(0.5) * step;
  const T sd = sqrt_t<T>(step);  // numpy sqrt(step) cast to dtype, mala.py:40 (correctly rounded in both)
  const uint32_t gchain = a.chain0 + (uint32_t)chain;
  uint32_t n_acc = 0;
  // per-chain dual-averaging tuner (HMC): every lane of the chain group carries the same state
  constexpr bool tuned = KIND == KIND_HMC_TUNED;
  double tn_barh = 0.0, tn_logbare = 0.0, tn_step = 0.0;
  int num_steps = a.num_steps;
  if (tuned) {
    tn_barh = a.tuner_state[chain];
    tn_logbare = a.tuner_state[a.n_chains + chain];
    tn_step = a.tuner_state[2 * a.n_chains + chain];
    num_steps = live ? (int)a.tuner_state[3 * a.n_chains + chain] : 1;   // padding groups do the minimum of work
    step = (T)tn_step;
    half_step = (T)(0.5 * tn_step);
  }

  for (long t = 0; t < a.n_iters; ++t) {
    T z[P];
#ifdef EB_THETA_IN_SMEM
    constexpr bool TSM = sizeof(T) == 8 && IS_HMC;
#else
    constexpr bool TSM = false;
#endif
    typename std::conditional<TSM, StridedVec<T>, RegVec<T, P>>::type thp;
    if constexpr (TSM) { thp.base = reinterpret_cast<T*>(smem + lay.off_grad) + threadIdx.x; thp.stride = kBlock; }
    typename std::conditional<GSM, StridedVec<T>, RegVec<T, P>>::type gp;
    if constexpr (GSM) { gp.base = reinterpret_cast<T*>(smem + lay.off_grad) + threadIdx.x; gp.stride = kBlock; }
    T u, ltp;
#ifndef EB_NO_PREFETCH_STATE   // measured on B200, config 4: 15.23e9 -> 15.40e9 evals/s
    // issue the loads of the chain's current state before the noise is generated (hundreds of cycles of arithmetic that do
    // not depend on them); hmc_draw reads the same addresses again and the compiler reuses the registers
    if constexpr (IS_HMC && !TSM && !GSM) {
#pragma unroll
      for (int j = 0; j < P; ++j) { thp[j] = cur.th[j * cur.stride]; gp[j] = cur.g[j * cur.stride]; }
    }
#endif
    if (a.rng_mode == 0) {
      philox_normals<T, P>(z, a.key, gchain, a.iter0 + (uint32_t)t);
      u = philox_uniform<T>(a.key, gchain, a.iter0 + (uint32_t)t);
    } else {
      const T* zt = a.z_tape + ((size_t)t * a.n_chains + chain) * P;
#pragma unroll
      for (int j = 0; j < P; ++j) z[j] = zt[j];
      u = a.u_tape[(size_t)t * a.n_chains + chain];
    }
    bool acc;
    if constexpr (KIND == KIND_MH) {
      acc = mh_draw<T, NET, G>(d, sub, step, a.symmetric != 0, cur, lt_cur, z, u, thp, ltp);
    } else if constexpr (KIND == KIND_MALA) {
      acc = mala_draw<T, NET, G>(d, sub, half_step, sd, cur, lt_cur, z, u, thp, gp, ltp);
    } else {
      T rate;
      // EB_MOM_IN_REGS: momentum as a register vector where one lane owns the chain.  Measured on B200 (config 4, 255
      // registers): 12.6e9 evals/s against 15.3e9 with the shared-memory column (the compiler already caches the column in
      // registers where it has room, and spills less) -> off
#ifdef EB_MOM_IN_REGS
      constexpr bool MOM_REGS = G == 1 && sizeof(T) == 8 && P <= 20 && !GSM && !TSM && min_blocks<T, NET, KIND>() <= 2;
#else
      constexpr bool MOM_REGS = false;
#endif
      if constexpr (MOM_REGS) {
        RegVec<T, P> pm;
        acc = hmc_draw<T, NET, G>(d, sub, step, half_step, num_steps, cur, lt_cur, z, pm, u, thp, gp, ltp, &rate);
      } else {
        StridedVec<T> pm{mom, CPB};
        acc = hmc_draw<T, NET, G>(d, sub, step, half_step, num_steps, cur, lt_cur, z, pm, u, thp, gp, ltp, &rate);
      }
      if (tuned && live && t < a.tuner_burnin) {                                          // hmc.py:158-163
        da_tune(a.tuner, (double)rate, a.tuner_iter0 + t + 1, t != a.tuner_burnin - 1, tn_barh, tn_logbare, tn_step,
                num_steps);
        step = (T)tn_step;
        half_step = (T)(0.5 * tn_step);
      }
    }
    if (acc) {  // uniform within the chain group: every lane holds identical values
      lt_cur = ltp;
      ++n_acc;
      if (live) {
#pragma unroll
        for (int j = 0; j < P; ++j) {
          if (j % G == sub) {
            cur.th[j * cur.stride] = thp[j];
            if (KIND != KIND_MH) cur.g[j * cur.stride] = gp[j];
          }
        }
      }
    }
    group_sync<G>();
    if (t >= a.n_burnin && (t - a.n_burnin) % a.thin == 0 && live) {  // serial_sampler.py:46
      const long s = (t - a.n_burnin) / a.thin;
      // the proposal registers hold the new state on accept; otherwise re-read the (unchanged) current state
      if (a.out_samples) {
#pragma unroll
        for (int j = 0; j < P; ++j)
          if (j % G == sub) a.out_samples[s * a.ss_i + chain * a.ss_c + j * a.ss_p] = acc ? thp[j] : cur.th[j * cur.stride];
      }
      if (KIND != KIND_MH && a.out_grad) {
#pragma unroll
        for (int j = 0; j < P; ++j)
          if (j % G == sub) a.out_grad[s * a.ss_i + chain * a.ss_c + j * a.ss_p] = acc ? gp[j] : cur.g[j * cur.stride];
      }
      if (sub == 0) {
        if (a.out_target) a.out_target[s * a.n_chains + chain] = lt_cur;
        if (a.out_acc) a.out_acc[s * a.n_chains + chain] = acc ? 1 : 0;
      }
    }
  }
  if (live && a.final_theta) {   // each lane re-reads the elements it owns (it wrote them itself on accept)
#pragma unroll
    for (int j = 0; j < P; ++j)
      if (j % G == sub) a.final_theta[chain * a.fs_c + j * a.fs_p] = cur.th[j * cur.stride];
  }
  if (live && sub == 0) {
    a.target[chain] = lt_cur;
    uint32_t total = n_acc;
    if (a.acc_count) { total += a.acc_count[chain]; a.acc_count[chain] = total; }
    if (a.final_target) a.final_target[chain] = lt_cur;
    if (a.final_acc) a.final_acc[chain] = total;
    if (tuned) {
      a.tuner_state[chain] = tn_barh;
      a.tuner_state[a.n_chains + chain] = tn_logbare;
      a.tuner_state[2 * a.n_chains + chain] = tn_step;
      a.tuner_state[3 * a.n_chains + chain] = (double)num_steps;
    }
  }
}

// ---- launchers ----------------------------------------------------------------------------------------------------
// Raises the kernel's dynamic shared-memory limit to `bytes`.  cudaErrorInvalidConfiguration when the staged data set
// (plus the kernel's static tables) exceeds what one CTA can have: capi.cu reports it as EEYORE_B200_EUNSUPPORTED.
template <class K> cudaError_t reserve_smem(K kern, size_t bytes) {
  int dev = 0, max_optin = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaFuncAttributes fa{};
  cudaError_t e = cudaFuncGetAttributes(&fa, kern);
  if (e != cudaSuccess) return e;
  if (bytes + fa.sharedSizeBytes > (size_t)max_optin) return cudaErrorInvalidConfiguration;
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

template <typename T, class NET, int G, int KIND> cudaError_t launch_sampler_g(const ChainArgs<T>& a, cudaStream_t st) {
  constexpr int KB = sampler_block<T, NET, KIND>();
  constexpr int CPB = KB / G;
#ifdef EB_THETA_IN_SMEM
  const SmemLayout<T, NET> lay(a.n_rows, CPB, KIND == KIND_HMC || KIND == KIND_HMC_TUNED, true);
#else
  const SmemLayout<T, NET> lay(a.n_rows, CPB, KIND == KIND_HMC || KIND == KIND_HMC_TUNED, grad_in_smem<T, NET, KIND>());
#endif
  auto kern = sampler_kernel<T, NET, G, KIND>;
  cudaError_t e = reserve_smem(kern, lay.total);
  if (e != cudaSuccess) return e;
  const long blocks = (a.n_chains + CPB - 1) / CPB;
  kern<<<(unsigned)blocks, KB, lay.total, st>>>(a);
  return cudaGetLastError();
}

template <typename T, class NET, int G> cudaError_t launch_eval_g(const ChainArgs<T>& a, cudaStream_t st) {
  constexpr int CPB = kBlock / G;
  const SmemLayout<T, NET> lay(a.n_rows, CPB, false);
  auto kern = eval_kernel<T, NET, G>;
  cudaError_t e = reserve_smem(kern, lay.total);
  if (e != cudaSuccess) return e;
  const long blocks = (a.n_chains + CPB - 1) / CPB;
  kern<<<(unsigned)blocks, kBlock, lay.total, st>>>(a);
  return cudaGetLastError();
}

#define EB_DISPATCH_G(G_, CALL)                 \
  switch (G_) {                                 \
    case 1: { constexpr int G = 1; CALL; }      \
    case 4: { constexpr int G = 4; CALL; }      \
    case 8: { constexpr int G = 8; CALL; }      \
    case 16: { constexpr int G = 16; CALL; }    \
    case 32: { constexpr int G = 32; CALL; }    \
    default: return cudaErrorInvalidValue;      \
  }

template <typename T, class NET> cudaError_t launch_sampler(int kind, int lanes, const ChainArgs<T>& a, cudaStream_t st) {
  switch (kind) {
    case KIND_MH: EB_DISPATCH_G(lanes, return (launch_sampler_g<T, NET, G, KIND_MH>(a, st)))
    case KIND_MALA: EB_DISPATCH_G(lanes, return (launch_sampler_g<T, NET, G, KIND_MALA>(a, st)))
    case KIND_HMC: EB_DISPATCH_G(lanes, return (launch_sampler_g<T, NET, G, KIND_HMC>(a, st)))
    case KIND_HMC_TUNED: EB_DISPATCH_G(lanes, return (launch_sampler_g<T, NET, G, KIND_HMC_TUNED>(a, st)))
  }
  return cudaErrorInvalidValue;
}

template <typename T, class NET> cudaError_t launch_eval(int lanes, const ChainArgs<T>& a, cudaStream_t st) {
  EB_DISPATCH_G(lanes, return (launch_eval_g<T, NET, G>(a, st)))
}

template <typename T, class NET> cudaError_t launch_forward(const ChainArgs<T>& a, T* out, cudaStream_t st) {
  const size_t sm = sizeof(T) * (size_t)a.n_rows * NET::D0;
  auto kern = forward_kernel<T, NET>;
  cudaError_t e = reserve_smem(kern, sm);
  if (e != cudaSuccess) return e;
  const long blocks = (a.n_chains + kBlock - 1) / kBlock;
  kern<<<(unsigned)blocks, kBlock, sm, st>>>(a, out);
  return cudaGetLastError();
}

}  // namespace eb
