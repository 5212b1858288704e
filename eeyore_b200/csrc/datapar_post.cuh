// The tail of one evaluation of the data-parallel path (BASELINE config 5), shared by the stand-alone post kernel
// (datapar.cu: dp_post_kernel) and the persistent HMC kernel (datapar_tc.cu: dp_hmc_run_kernel):
//   1. fixed-order sum of the per-CTA partial sums of the evaluation
//   2. the exchange step of the data-sharded path: every rank stores its 1 + P sums straight into every peer's inbox over
//      NVLink (peer pointers from CUDA IPC), releases a per-CTA flag with the evaluation's sequence number, waits for the
//      same flag from every peer and adds the W inbox slots in rank order -> bit-identical totals on every rank
//      (was an NCCL all-reduce of 42.5 KB, latency-bound, plus a launch gap on either side)
//   3. Normal log-prior and its gradient, temperature (eeyore/models/bayesian_model.py:46-56)
//   4. optionally the leapfrog update that follows the evaluation (eeyore/samplers/hmc.py:113-119)
// Layout: 21 CTAs x 256 threads, one thread per entry of [loglik, dloglik].
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "datapar.cuh"

namespace eb {

constexpr double kDpLogSqrt2Pi = 0.9189385332046727;     // log(sqrt(2 pi))
constexpr int DP_POST_THREADS = 256;
constexpr int DP_POST_CTAS = (DP_P + 1 + DP_POST_THREADS - 1) / DP_POST_THREADS;   // 21
constexpr int DP_XSLOT = DP_POST_CTAS * DP_POST_THREADS;                          // doubles per inbox slot (5376)
constexpr int DP_XMAXW = 8;                                                        // ranks per box
// scratch (DP_SCRATCH_LEN doubles): [0..31] log-prior partials, [32..63] kinetic partials, [64] summed log-likelihood,
// [65] ticket counter (as u64; stand-alone post kernel only), [128..191] kinetic partials of the momentum draw, double
// buffered by iteration parity (persistent kernel only: a CTA may start the next iteration's draw while another still reads
// this iteration's partials in its accept test)
constexpr int DP_SCRATCH_LP = 0, DP_SCRATCH_KIN = 32, DP_SCRATCH_LL = 64, DP_SCRATCH_TICKET = 65, DP_SCRATCH_KIN0 = 128;
constexpr int DP_SCRATCH_LEN = 256;

struct DpExchange {
  int world, rank;
  unsigned long long seq;                 // evaluation counter, identical on every rank, starts at 1
  double* inbox[DP_XMAXW];                // inbox[p]: rank p's [2 parity][8 src][DP_XSLOT] doubles (peer-mapped)
  unsigned long long* flags[DP_XMAXW];    // flags[p]: rank p's [2 parity][8 src][32 cta] sequence numbers
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// barrier over the DP_POST_THREADS post / epilogue threads (named barrier 5).  In the stand-alone post kernel that is the whole
// CTA; in the persistent run kernel the CTA also holds the MMA-issue warp group, which never takes part in these.
__device__ __forceinline__ void post_sync() { asm volatile("bar.sync 5, %0;" ::"n"(DP_POST_THREADS) : "memory"); }

__device__ __forceinline__ double block_sum_256(double v, double* red) {
  // fixed order: lanes by xor-shuffle, then the 8 warp sums in warp order
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  post_sync();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < DP_POST_THREADS / 32; ++w) t += red[w];
  post_sync();
  return t;   // valid in thread 0
}

// Thread (cta, tid) of the post layout; every thread of the CTA calls it (CTA barriers inside).  Writes the gradient (and,
// with step_mode 1 / 2, the inner / last leapfrog update) of its entry, and the CTA's log-prior / kinetic partial sums to
// scratch; entry 0 leaves the summed log-likelihood in scratch[DP_SCRATCH_LL].  `red`: 8 doubles of shared memory.
// All reads of data written by other CTAs of the same launch (partials) or by peers (inbox) bypass L1.
__device__ __forceinline__ void dp_post_entries(const double* partials, int n_parts, const DpExchange& xc, unsigned long long seq,
                                                const float* theta, const float* ploc, const float* pscale, int has_temp,
                                                double temp, float* grad_out, int step_mode, float step, float* mom,
                                                float* theta_p, double* scratch, int* status, double* red, int cta, int tid) {
  const int e = cta * DP_POST_THREADS + tid;          // entry of [loglik, dloglik]; e <= DP_P is valid
  double tot = 0.0;
  if (e <= DP_P) {
    // fixed order c = 0, 1, ...; the loads of 16 rows are in flight together (one dependent load per add would make the fold
    // a chain of n_parts L2 round trips: it was the largest fixed cost of an evaluation)
    const double* col = partials + e;
    int c = 0;
    for (; c + 16 <= n_parts; c += 16) {
      double v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __ldcg(col + (size_t)(c + i) * (DP_P + 1));
#pragma unroll
      for (int i = 0; i < 16; ++i) tot += v[i];
    }
    for (; c < n_parts; ++c) tot += __ldcg(col + (size_t)c * (DP_P + 1));
  }
  if (xc.world > 1) {
    const int par = (int)(seq & 1ull);
    const size_t slot = (size_t)(par * DP_XMAXW + xc.rank) * DP_XSLOT + e;
    for (int p = 0; p < xc.world; ++p) xc.inbox[p][slot] = tot;      // coalesced 2 KB per CTA per peer, over NVLink
    __threadfence_system();
    post_sync();
    // thread p < world releases this CTA's flag on peer p and waits for peer p's flag here: the W system-scope stores and the
    // W polling loops run side by side (one thread doing them in turn cost a round trip per peer)
    if (tid < xc.world) {
      st_release_sys(&xc.flags[tid][(par * DP_XMAXW + xc.rank) * 32 + cta], seq);
      const unsigned long long* f = &xc.flags[xc.rank][(par * DP_XMAXW + tid) * 32 + cta];
      long spins = 0;
      while (ld_acquire_sys(f) < seq) {
        if (++spins > (1L << 28)) {   // seconds: a peer is gone; fail loudly instead of hanging the box
          status[0] = 1;
          __trap();
        }
      }
    }
    post_sync();
    tot = 0.0;
    for (int src = 0; src < xc.world; ++src)
      tot += __ldcg(&xc.inbox[xc.rank][(size_t)(par * DP_XMAXW + src) * DP_XSLOT + e]);
  }
  double lp = 0.0, kin = 0.0;
  if (e == 0) scratch[DP_SCRATCH_LL] = tot;
  if (e >= 1 && e <= DP_P) {
    const int j = e - 1;
    const double sc = (double)pscale[j], dd = (double)theta[j] - (double)ploc[j];
    lp = -(dd * dd) / (2.0 * sc * sc) - log(sc) - kDpLogSqrt2Pi;
    double g = tot - dd / (sc * sc);
    if (has_temp) g *= temp;
    const float gf = (float)g;
    grad_out[j] = gf;
    if (step_mode) {
      const float w = (step_mode == 2) ? 0.5f * step : step;
      const float pj = fmaf(w, gf, mom[j]);
      mom[j] = pj;
      if (step_mode == 1) theta_p[j] = fmaf(step, pj, theta_p[j]);
      kin = (double)pj * (double)pj;
    }
  }
  const double lp_cta = block_sum_256(lp, red);
  const double kin_cta = block_sum_256(kin, red);
  if (tid == 0) {
    scratch[DP_SCRATCH_LP + cta] = lp_cta;
    scratch[DP_SCRATCH_KIN + cta] = kin_cta;
  }
}

// target = T (loglik + sum of the CTA log-prior partials), kinetic energy = sum of the CTA partials / 2, CTA order
__device__ __forceinline__ void dp_post_scalars(const double* scratch, int has_temp, double temp, double& target, double& kin) {
  double lpt = 0.0, kt = 0.0;
  for (int c = 0; c < DP_POST_CTAS; ++c) {
    lpt += __ldcg(&scratch[DP_SCRATCH_LP + c]);
    kt += __ldcg(&scratch[DP_SCRATCH_KIN + c]);
  }
  double ll = __ldcg(&scratch[DP_SCRATCH_LL]);
  if (has_temp) { ll *= temp; lpt *= temp; }
  target = ll + lpt;
  kin = 0.5 * kt;
}

}  // namespace eb
