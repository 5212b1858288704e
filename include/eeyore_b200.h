/*
 * eeyore_b200 -- C ABI of the B200-native sampler inner loop.
 *
 * The reference (papamarkou/eeyore) is pure Python and has no FFI; the seam this library sits behind is the
 * Python method surface listed in SURVEY.md section 8(b).  Each entry point below cites the reference method
 * it replaces (paths relative to the reference root).  INTEGRATION.md shows the ctypes binding a maintainer
 * of the reference would add.
 *
 * Conventions
 *   - every pointer is DEVICE-ADDRESSABLE memory unless its name ends in _host; all work is enqueued on `stream`
 *     (a cudaStream_t passed as void*, NULL = legacy default stream) and is asynchronous.  The saved-state outputs of the
 *     run entry points (out_samples / out_grad / out_target / out_accepted) may also be PINNED HOST memory
 *     (cudaHostAlloc; directly addressable by the device under unified addressing): the kernel's coalesced stores then
 *     are the device->host transfer of the chain (what ChainList holds in the reference, chains/chain_list.py:12-30),
 *     overlapped with the sampling; contents are valid once `stream` has been synchronised;
 *   - dtype: EEYORE_B200_F32 or EEYORE_B200_F64 -- theta, x, y, prior and outputs all use it
 *     (the reference's model.dtype, eeyore/models/model.py:7);
 *   - theta is the reference's flat parameter vector (eeyore/models/model.py:44-55): per layer the weight
 *     matrix row-major [d_out, d_in] followed by the bias [d_out]; C chains are stored [C, P] row-major;
 *   - x is [N, d_0] row-major; y is [N] (binary, values in {0,1}) or one-hot [N, K] (multiclass), as produced
 *     by eeyore/datasets/xydataset.py:11-53;
 *   - return value: 0 on success, a negative EEYORE_B200_E* code otherwise; eeyore_b200_last_error() gives
 *     the message (thread-local).  The Python layer raises RuntimeError / ValueError from it, matching the
 *     reference's error behaviour.
 */
#ifndef EEYORE_B200_H
#define EEYORE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EEYORE_B200_F32 0
#define EEYORE_B200_F64 1

#define EEYORE_B200_ACT_NONE 0    /* activations[l] is None          (eeyore/models/mlp.py:48) */
#define EEYORE_B200_ACT_SIGMOID 1 /* activations[l] is torch.sigmoid (eeyore/models/mlp.py:48-49) */

#define EEYORE_B200_LOSS_BINARY 0     /* loss_functions['binary_classification']     (eeyore/constants/constants.py:16) */
#define EEYORE_B200_LOSS_MULTICLASS 1 /* loss_functions['multiclass_classification'] (eeyore/constants/constants.py:17) */

#define EEYORE_B200_RNG_PHILOX 0 /* on-device Philox4x32-10 keyed by (seed, chain, iteration) */
#define EEYORE_B200_RNG_TAPE 1   /* noise read from z_tape / u_tape (parity runs against the reference) */

#define EEYORE_B200_OK 0
#define EEYORE_B200_EINVAL (-1)      /* bad argument           -> ValueError   */
#define EEYORE_B200_EUNSUPPORTED (-2)/* architecture not built -> ValueError   */
#define EEYORE_B200_ECUDA (-3)       /* CUDA runtime failure   -> RuntimeError */
#define EEYORE_B200_ENUMERIC (-4)    /* numerical failure      -> RuntimeError */

typedef struct eeyore_b200_mlp *eeyore_b200_mlp_t;

const char *eeyore_b200_last_error(void);
const char *eeyore_b200_version(void);

/* mlp.Hyperparameters + MLP.__init__ (eeyore/models/mlp.py:9-43): describe the network once.
 * dims[n_layers+1]; bias[n_layers] (0/1); act_ids[n_layers]; loss_id; dtype. */
int eeyore_b200_mlp_create(int n_layers, const int *dims, const int *bias, const int *act_ids, int loss_id,
                           int dtype, eeyore_b200_mlp_t *out);
int eeyore_b200_mlp_destroy(eeyore_b200_mlp_t h);
/* Model.num_params (eeyore/models/model.py:34-36) */
int eeyore_b200_mlp_num_params(eeyore_b200_mlp_t h);
/* 1 if the network is served by a compile-time specialisation (2-2-1, 2-3-2-1, 4-3-3, 4-3-2-3 with all biases and
 * sigmoid hidden units), 0 if by the runtime-shape kernels (any dims <= 8 layers, bias flags, sigmoid / None) */
int eeyore_b200_mlp_is_specialised(eeyore_b200_mlp_t h);

/* BayesianModel.log_target + LogTargetModel.upto_grad_log_target
 * (eeyore/models/bayesian_model.py:52-56, eeyore/models/log_target_model.py:20-23), batched over C chains.
 * out_grad may be NULL (log_target only).  out_loglik / out_logprior may be NULL
 * (BayesianModel.log_lik / log_prior, bayesian_model.py:30-35,46-50).
 * has_temperature = 0 reproduces temperature=None. */
int eeyore_b200_log_target_grad(eeyore_b200_mlp_t h, int64_t n_chains, const void *theta, const void *x,
                                const void *y, int64_t n_rows, const void *prior_loc, const void *prior_scale,
                                int has_temperature, double temperature, void *out_target, void *out_grad,
                                void *out_loglik, void *out_logprior, int lanes_per_chain, void *stream);

/* MLP.forward (eeyore/models/mlp.py:45-50): out [C, N, d_L] (probabilities for a sigmoid head, logits otherwise) */
int eeyore_b200_forward(eeyore_b200_mlp_t h, int64_t n_chains, const void *theta, const void *x, int64_t n_rows,
                        void *out, void *stream);

typedef struct eeyore_b200_run_params {
  int64_t n_chains;        /* C independent chains (SerialSampler.benchmark semantics, serial_sampler.py:54-126) */
  int64_t n_iters;         /* iterations run by this call (= num_epochs when num_batches == 1) */
  int64_t n_burnin;        /* leading iterations whose state is not saved (serial_sampler.py:46) */
  int64_t thin;            /* keep every thin-th post-burn-in state; 1 = reference behaviour */
  double step;             /* MALA/SMMALA/HMC step (mala.py:12, hmc.py:11); MH: proposal scale (metropolis_hastings.py:27) */
  int32_t num_steps;       /* HMC leapfrog steps (hmc.py:11) */
  int32_t symmetric;       /* MH: 1 = symmetric proposal (metropolis_hastings.py:10) */
  int32_t has_temperature; /* 0 = temperature None */
  int32_t rng_mode;        /* EEYORE_B200_RNG_* */
  double temperature;
  uint64_t seed;           /* Philox key */
  uint64_t iter_offset;    /* Philox counter word: global iteration index of the first iteration of this call */
  uint64_t chain_offset;   /* global chain id of local chain 0 (chain sharding across GPUs) */
  const void *z_tape;      /* [n_iters, C, P] standard normals (tape mode) */
  const void *u_tape;      /* [n_iters, C] uniforms (tape mode) */
  const void *x;           /* [N, d0] */
  const void *y;           /* [N] or [N, K] */
  int64_t n_rows;
  const void *prior_loc;   /* [P] */
  const void *prior_scale; /* [P] */
  void *theta;             /* in/out current['sample']     [C, P] (see st_chain / st_param) */
  void *target;            /* in/out current['target_val'] [C]    */
  void *grad;              /* in/out current['grad_val']   [C, P] (unused by MH) */
  int64_t st_chain, st_param; /* layout of theta / grad: element (c, j) at c*st_chain + j*st_param; 0,0 = row-major
                                 [C, P] (st_chain = P, st_param = 1).  The device-resident path uses the chain-minor
                                 layout (st_chain = 1, st_param = C) so that warps read and write the state coalesced. */
  void *out_samples;       /* saved 'sample' states, element (s,c,j) at s*ss_iter + c*ss_chain + j*ss_param; may be NULL */
  int64_t ss_iter, ss_chain, ss_param;
  void *out_target;        /* [n_saved, C] saved 'target_val'; may be NULL */
  void *out_grad;          /* saved 'grad_val', same strides as out_samples; may be NULL */
  uint8_t *out_accepted;   /* [n_saved, C] saved 'accepted'; may be NULL */
  uint32_t *accept_count;  /* [C], incremented by the number of accepted proposals over all n_iters; may be NULL */
  int32_t lanes_per_chain; /* threads cooperating on one chain: 1, 4, 8, 16 or 32; 0 = auto (else EINVAL) */
  int32_t reserved;
  /* HMC dual-averaging step-size tuner: HMCDATuner.tune hooked into HMC.draw during burn-in
   * (eeyore/tuners/hmcda_tuner.py:8-59, eeyore/samplers/hmc.py:158-163).  tuner_state == NULL disables it.
   * One independent tuner per chain; python-float (fp64) arithmetic as in the reference. */
  double tuner_l;          /* target trajectory length: num_steps = max(1, round(l / step)) */
  double tuner_d;          /* target acceptance rate (0.65) */
  double tuner_m;          /* log(10 * e0), HMCDATuner.set_m */
  double tuner_logeub;     /* log of the step upper bound (used when tuner_has_eub) */
  int32_t tuner_has_eub;
  int32_t tuner_pad;
  int64_t tuner_iter0;     /* counter.idx at the first iteration of this call */
  int64_t tuner_burnin;    /* leading iterations of this call that still tune (num_burnin_iters - counter.idx, >= 0) */
  double *tuner_state;     /* in/out [4, C]: barh, logbare, step, num_steps */
  void *stream;
  /* Adaptive random-walk samplers (am_run / ram_run only).
   * AM  (eeyore/samplers/am.py:8-107):  adapt_p = {l, b, c}, adapt_t0 = t0, adapt_cov0 = cov0 [P, P];
   *      adapt_state [C, P + 2 P^2 + 1] = running mean, sum of theta theta^T, covariance estimate, number of accepted moves
   *      (zeros / zeros / cov0 / 0 at the start of a chain); u_tape holds TWO uniforms per iteration [n_iters, 2, C].
   * RAM (eeyore/samplers/ram.py:7-70):  adapt_p = {a, g, -};  adapt_state [C, P^2] = lower Cholesky factor of the proposal.
   * adapt_iter0 = counter.idx at the first iteration of the call.  adapt_status [C]: 0, or 1 + iteration at which the
   * factorisation failed (the reference's torch.linalg.cholesky raises there); the chain stops moving from then on. */
  double adapt_p[3];
  int32_t adapt_t0;
  int32_t adapt_pad;
  int64_t adapt_iter0;
  void *adapt_state;
  const void *adapt_cov0;
  int32_t *adapt_status;
  /* Final state of the call written a second time, for the host (mh_run / mala_run / hmc_run on the compiled network
   * specialisations; ignored elsewhere; NULL = off).  Like the saved-state outputs these may be pinned host memory: every
   * chain's last stores are then its device->host transfer, spread over the run as the CTAs finish, instead of a copy
   * after the kernel.  final_theta: element (c, j) at c*fs_chain + j*fs_param (chain-minor -- fs_chain = 1, fs_param = C
   * -- keeps a warp's stores contiguous); final_target [C]; final_accept_count [C] = accept_count after this call (the
   * number accepted in this call when accept_count is NULL). */
  void *final_theta;
  int64_t fs_chain, fs_param;
  void *final_target;
  uint32_t *final_accept_count;
} eeyore_b200_run_params;

/* number of states a run with these (n_iters, n_burnin, thin) saves */
int64_t eeyore_b200_num_saved(int64_t n_iters, int64_t n_burnin, int64_t thin);

/* MetropolisHastings.draw x n_iters (eeyore/samplers/metropolis_hastings.py:41-73) */
int eeyore_b200_mh_run(eeyore_b200_mlp_t h, const eeyore_b200_run_params *p);
/* AM.draw x n_iters (eeyore/samplers/am.py:62-107) and RAM.draw x n_iters (eeyore/samplers/ram.py:39-70): one warp per
 * chain, proposal factor adapted and re-factorised in shared memory; compiled network specialisations with P <= 32 */
int eeyore_b200_am_run(eeyore_b200_mlp_t h, const eeyore_b200_run_params *p);
int eeyore_b200_ram_run(eeyore_b200_mlp_t h, const eeyore_b200_run_params *p);
/* length (in elements of the model dtype) of one chain's adapt_state: kind 0 = AM, 1 = RAM */
int64_t eeyore_b200_adapt_state_len(eeyore_b200_mlp_t h, int kind);
/* MALA.draw x n_iters (eeyore/samplers/mala.py:46-82) */
int eeyore_b200_mala_run(eeyore_b200_mlp_t h, const eeyore_b200_run_params *p);
/* HMC.draw + HMC.leapfrog x n_iters (eeyore/samplers/hmc.py:100-170) */
int eeyore_b200_hmc_run(eeyore_b200_mlp_t h, const eeyore_b200_run_params *p);
/* SMMALA (absent from the reference snapshot; SURVEY.md A.7, builder-defined): Fisher metric, batched Cholesky */
int eeyore_b200_smmala_run(eeyore_b200_mlp_t h, const eeyore_b200_run_params *p);

/* Chain diagnostics on the device, one warp per chain (samples element (s, c, j) at s*ss_iter + c*ss_chain + j*ss_param;
 * n_params <= 32).  Any output pointer may be NULL.
 *   out_mean [C,P]     ChainList.mean                        (eeyore/chains/chain_list.py:69-71)
 *   out_cov  [C,P,P]   stats.cov                             (eeyore/stats/cov.py:5-15)
 *   out_inse [C,P,P]   stats.inse_mc_cov, adjust=False       (eeyore/stats/inse_mc_cov.py:9-83)
 *   out_ess  [C]       stats.multi_ess                       (eeyore/stats/multi_ess.py:6-14)
 *   out_status [C]     0 = ok, 1 = 'Not enough samples' (inse_mc_cov.py:44-45; the Python layer raises RuntimeError),
 *                      3 = undecided after defer_after lag pairs (only when defer_after >= 0)
 *   chain_index        NULL, or n_chains device indices: only these chains are processed (inputs and outputs are addressed
 *                      by the index).  A chain whose estimate never becomes positive definite walks through n / 2 lag
 *                      pairs; with defer_after = m >= 0 such chains stop after lag pair m with status 3 and the caller
 *                      finishes them in a second call (chain_index = those chains, defer_after = -1), where they all run side
 *                      by side.  defer_after = -1: every chain is finished in this call.
 *   out_lags [C,2]     (first lag index with a positive-definite estimate, last accepted lag index)
 *   out_acf  [C, max_lag+1, P]  autocorrelation function (builder-defined; absent from the reference) */
int eeyore_b200_chain_stats(int dtype, int64_t n_chains, int64_t n_samples, int n_params, const void *samples,
                            int64_t ss_iter, int64_t ss_chain, int64_t ss_param, void *out_mean, void *out_cov,
                            void *out_inse, void *out_ess, int32_t *out_status, int32_t *out_lags, int max_lag,
                            void *out_acf, const int64_t *chain_index, int defer_after, void *stream);

/* ---- data-parallel path (BASELINE config 5: MLP 16-64-64-1, fp32, rows sharded across GPUs) -----------------------
 * One parameter vector, millions of rows.  Each rank evaluates its row shard; the caller all-reduces out_sums
 * (P + 1 doubles) over the ranks (NCCL) and then calls dp_finish, which adds the prior once.
 * Replaces eeyore/models/log_target_model.py:20-23 for large N; theta fp32 [P] in the reference layout. */
int eeyore_b200_dp_num_params(void);
/* out_sums[0] = sum_i loglik_i, out_sums[1 + j] = d/dtheta_j sum_i loglik_i over this rank's rows (fp64, deterministic) */
int eeyore_b200_dp_loglik_grad(const void *theta, const void *x, const void *y, int64_t n_rows, void *out_sums,
                               void *workspace, void *stream);
/* the same with the shard's max |x| supplied as a device scalar (fp32, e.g. from dp_absmax, computed once per data set):
 * the tcgen05 kernel scales x by a power of two before splitting it into fp16 pieces; x_absmax NULL = computed per call */
int eeyore_b200_dp_loglik_grad_x(const void *theta, const void *x, const void *y, int64_t n_rows, const void *x_absmax,
                                 void *out_sums, void *workspace, void *stream);
int eeyore_b200_dp_absmax(const void *x, int64_t n_values, void *out_absmax, void *stream);
/* out_sums may be NULL when a workspace is given: the per-CTA rows ([dp_num_parts(n_rows)][P + 1] doubles) are then left
 * in the workspace for dp_post, which folds them together with the exchange step. */
int eeyore_b200_dp_num_parts(int64_t n_rows);
/* the same sums from the FP32 CUDA-core formulation (no tensor cores): kept as an independent cross-check of the
 * tcgen05 kernel and as the A/B line of bench.py; not used by the product path */
int eeyore_b200_dp_loglik_grad_ffma(const void *theta, const void *x, const void *y, int64_t n_rows, void *out_sums,
                                    void *workspace, void *stream);
/* size of the optional caller-owned workspace of dp_loglik_grad (per-CTA partial sums); NULL workspace = temporary */
int64_t eeyore_b200_dp_workspace_bytes(void);
/* target (fp64 scalar) and gradient (fp32 [P]) from the all-reduced sums: adds the Normal log-prior
 * (eeyore/models/bayesian_model.py:46-56) and applies the temperature */
int eeyore_b200_dp_finish(const void *sums, const void *theta, const void *prior_loc, const void *prior_scale,
                          int has_temperature, double temperature, void *out_target, void *out_grad, void *stream);
/* Fused tail of one evaluation (one launch): fixed-order fold of the workspace rows; for world > 1 the exchange step of
 * the data-sharded path -- every rank stores its 1 + P sums into every peer's exchange area over NVLink (peer_bases[p] =
 * rank p's area, mapped with dp_exchange_open), releases sequence-numbered flags and adds the world slots in rank order,
 * so every rank holds bit-identical totals without NCCL; then the Normal log-prior, its gradient and the temperature
 * (eeyore/models/bayesian_model.py:46-56, log_target_model.py:20-23) and, for step_mode 1 (inner) / 2 (last), the
 * leapfrog update that follows the evaluation (eeyore/samplers/hmc.py:113-119).  seq = evaluation counter (>= 1, the same
 * on every rank).  local_scratch = this rank's scratch (dp_exchange_scratch).  status[0] != 0 after a peer time-out. */
int eeyore_b200_dp_post(const void *workspace, int n_parts, int world, int rank, uint64_t seq, void *const *peer_bases,
                        void *local_scratch, const void *theta, const void *prior_loc, const void *prior_scale,
                        int has_temperature, double temperature, void *out_grad, void *out_target, int step_mode,
                        double step, void *momentum, void *theta_prop, void *kin1, int32_t *status, void *stream);
/* exchange area of one rank: create (cudaMalloc, zeroed) + its 64-byte CUDA IPC handle; open / close a peer's area */
int64_t eeyore_b200_dp_exchange_bytes(void);
int eeyore_b200_dp_exchange_create(void **out_base, void *out_handle64);
int eeyore_b200_dp_exchange_open(const void *handle64, void **out_ptr);
int eeyore_b200_dp_exchange_close(void *peer_ptr);
int eeyore_b200_dp_exchange_destroy(void *base);
/* byte offset of the scratch block (dp_scratch_len doubles) inside an exchange area */
int64_t eeyore_b200_dp_exchange_scratch_offset(void);
/* doubles a local scratch block must hold (zero-initialised by the caller) */
int64_t eeyore_b200_dp_scratch_len(void);
/* HMC.draw pieces for a replicated chain state (eeyore/samplers/hmc.py:100-170): momentum draw + first half step;
 * momentum / position update after each evaluation; accept test and commit.  z_tape / u_tape NULL = Philox. */
int eeyore_b200_dp_hmc_begin(const void *theta_cur, const void *grad_cur, double step, uint64_t seed, uint64_t iter,
                             const void *z_tape, void *momentum, void *theta_prop, void *kin0, void *stream);
int eeyore_b200_dp_hmc_step(const void *grad_prop, double step, int last, void *momentum, void *theta_prop, void *kin1,
                            void *stream);
int eeyore_b200_dp_hmc_accept(void *theta_cur, void *grad_cur, void *target_cur, const void *theta_prop,
                              const void *grad_prop, const void *target_prop, const void *kin0, const void *kin1,
                              uint64_t seed, uint64_t iter, const void *u_tape, void *out_sample, void *out_target,
                              uint8_t *out_accepted, uint32_t *accept_count, void *stream);

/* sampler.run for the replicated chain of the data-sharded path in ONE cooperative launch (persistent CTAs, one per SM):
 * n_iters x [HMC.draw: momentum draw, num_steps x (evaluation over this rank's rows on the tensor cores, grid barrier, fold +
 * peer-store exchange + prior + leapfrog, grid barrier), accept test, commit, sample write-out]
 * (eeyore/samplers/hmc.py:100-170 inside the loop of eeyore/samplers/serial_sampler.py:35-52).  theta_cur / grad_cur /
 * target_cur (fp64 scalar) hold the current state on entry (dp_loglik_grad_x + dp_post at theta_cur) and on return.
 * workspace: dp_workspace_bytes; local_scratch: dp_scratch_len doubles; grid_counter: 8 bytes of device memory.
 * The first evaluation uses exchange sequence number `seq`, the k-th seq + k: the caller advances its counter by
 * n_iters * num_steps.  Saved from iteration n_burnin on: out_samples [n_saved, P] fp32, out_target [n_saved] fp64,
 * out_accepted [n_saved]; each may be NULL.  z_tape [n_iters, P] / u_tape [n_iters] NULL = Philox.
 * status[0]: 1 = a peer never delivered its sums, 2 = grid barrier time-out (the kernel traps). */
int eeyore_b200_dp_hmc_run(const void *x, const void *y, int64_t n_rows, const void *x_absmax, void *theta_cur, void *grad_cur,
                           void *target_cur, void *theta_prop, void *grad_prop, void *momentum, void *workspace,
                           void *local_scratch, void *grid_counter, int32_t *status, const void *prior_loc,
                           const void *prior_scale, int has_temperature, double temperature, double step, int num_steps,
                           int64_t n_iters, int64_t n_burnin, uint64_t seed, uint64_t iter_offset, const void *z_tape,
                           const void *u_tape, void *out_samples, void *out_target, uint8_t *out_accepted,
                           uint32_t *accept_count, int world, int rank, uint64_t seq, void *const *peer_bases, void *stream);

/* Philox draws exactly as the samplers consume them (tests / reproducibility):
 * out_z [n_chains, P] normals and out_u [n_chains] uniform of iteration `iter`. */
int eeyore_b200_philox_draws(int dtype, int64_t n_chains, int n_params, uint64_t seed, uint64_t iter,
                             uint64_t chain_offset, void *out_z, void *out_u, void *stream);

/* Measured FMA peak of the device (dependent-free FFMA/DFMA chains); roofline denominator for the chain kernels */
int eeyore_b200_fma_peak(int dtype, int iters, double *out_tflops);

#ifdef __cplusplus
}
#endif
#endif /* EEYORE_B200_H */
