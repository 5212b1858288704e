"""Oracle (test infrastructure): MH / MALA / HMC / SMMALA draws, batched over chains, driven by a noise tape.

numpy restatement of
  eeyore/samplers/metropolis_hastings.py:25-73  (random-walk MH, log-space accept)
  eeyore/samplers/mala.py:35-82                 (MALA; q = N(theta + step/2 grad, step I); log-space accept)
  eeyore/samplers/hmc.py:91-170                 (leapfrog, hamiltonian, linear-space accept)
  eeyore/kernels/normalized_kernel.py:14-19, normal_kernel.py:5-23  (Normal sample / summed log_prob)
  eeyore/samplers/serial_sampler.py:35-52       (burn-in gating of saved states)
for the full-batch case (num_batches == 1: cached current target/grad are reused, mala.py:49,68).

The noise tape replaces the reference's global torch RNG: ``z[t, c, :]`` is the
standard-normal draw of iteration t (proposal noise for MH/MALA/SMMALA, initial
momentum for HMC) and ``u[t, c]`` the uniform of the accept test, in the
reference's per-iteration order (z first, then u).

SMMALA is builder-defined: the reference snapshot has no SMMALA sampler and the restatement follows SURVEY.md A.7.  It
is pinned by tests/golden/smmala_*.npz -- runs that oracle/make_golden.py:smmala_goldens assembles from the reference's
own pieces only (MLP + autograd row derivatives, is_pos_def, MultivariateNormalKernel.log_prob, MALA's accept rule).
"""
from __future__ import annotations

import math

import numpy as np

from .mlp import MLPSpec, log_target, log_target_grad

_HALF_LOG_2PI = math.log(math.sqrt(2 * math.pi))


def _normal_log_prob_sum(value, loc, scale):
    """torch.distributions.Normal.log_prob summed over parameters (normalized_kernel.py:15)."""
    dt = value.dtype
    var = scale * scale
    return (-((value - loc) ** 2) / (2 * var) - np.log(scale) - dt.type(_HALF_LOG_2PI)).sum(axis=-1)


def _collect(store, t, n_burnin, thin, **state):
    if t >= n_burnin and (t - n_burnin) % thin == 0:
        for k, v in state.items():
            store.setdefault(k, []).append(np.array(v, copy=True))


def _finish(store, final):
    out = {k: np.stack(v) for k, v in store.items()}
    out["final"] = final
    return out


def mh_run(spec: MLPSpec, x, y, loc, scale, theta0, z, u, n_burnin=0, prop_scale=1.0,
           temperature=None, thin=1, symmetric=True):
    """metropolis_hastings.py:41-73.  theta0 [C,P]; z [T,C,P]; u [T,C]."""
    theta = np.array(np.atleast_2d(theta0), copy=True)
    dt = theta.dtype
    ps = np.full(theta.shape[1], prop_scale, dtype=dt)
    lt = log_target(spec, theta, x, y, loc, scale, temperature)
    store = {}
    for t in range(z.shape[0]):
        prop = theta + ps * z[t].astype(dt)                       # kernel.sample(), loc = current sample
        lt_p = log_target(spec, prop, x, y, loc, scale, temperature)
        log_rate = lt_p - lt
        if not symmetric:                                          # :51-54
            log_rate = log_rate - _normal_log_prob_sum(prop, theta, ps)
            log_rate = log_rate + _normal_log_prob_sum(theta, prop, ps)
        with np.errstate(divide="ignore", invalid="ignore"):
            acc = np.log(u[t].astype(dt)) < log_rate              # :56, NaN compares False
        theta = np.where(acc[:, None], prop, theta)
        lt = np.where(acc, lt_p, lt)
        _collect(store, t, n_burnin, thin, sample=theta, target_val=lt, accepted=acc.astype(np.uint8))
    return _finish(store, dict(sample=theta, target_val=lt))


def mala_run(spec: MLPSpec, x, y, loc, scale, theta0, z, u, step, n_burnin=0, temperature=None, thin=1):
    """mala.py:46-82."""
    theta = np.array(np.atleast_2d(theta0), copy=True)
    dt = theta.dtype
    half_step = dt.type(0.5 * step)                                # python-float product, mala.py:36
    ps = np.full(theta.shape[1], np.sqrt(step), dtype=dt)          # mala.py:40
    lt, g = log_target_grad(spec, theta, x, y, loc, scale, temperature)
    store = {}
    for t in range(z.shape[0]):
        mean_c = theta + half_step * g                             # kernel_mean(current)
        prop = mean_c + ps * z[t].astype(dt)                       # :53
        lt_p, g_p = log_target_grad(spec, prop, x, y, loc, scale, temperature)
        with np.errstate(invalid="ignore"):
            log_rate = lt_p - lt                                   # :58
            log_rate = log_rate - _normal_log_prob_sum(prop, mean_c, ps)          # :60
            mean_p = prop + half_step * g_p                        # :62
            log_rate = log_rate + _normal_log_prob_sum(theta, mean_p, ps)         # :64
            acc = np.log(u[t].astype(dt)) < log_rate               # :66
        theta = np.where(acc[:, None], prop, theta)
        g = np.where(acc[:, None], g_p, g)
        lt = np.where(acc, lt_p, lt)
        _collect(store, t, n_burnin, thin, sample=theta, target_val=lt, grad_val=g,
                 accepted=acc.astype(np.uint8))
    return _finish(store, dict(sample=theta, target_val=lt, grad_val=g))


def leapfrog(spec: MLPSpec, x, y, loc, scale, theta, p0, step, num_steps, temperature=None):
    """hmc.py:100-124 (identity mass; num_steps+1 gradient evaluations; final momentum negation)."""
    dt = theta.dtype
    eps = dt.type(step)
    half = dt.type(0.5 * step)
    pos = np.array(theta, copy=True)
    with np.errstate(invalid="ignore", over="ignore"):
        lt, g = log_target_grad(spec, pos, x, y, loc, scale, temperature)
        mom = p0 - half * (-g)
        for _ in range(num_steps - 1):
            pos = pos + eps * mom
            lt, g = log_target_grad(spec, pos, x, y, loc, scale, temperature)
            mom = mom - eps * (-g)
        pos = pos + eps * mom
        lt, g = log_target_grad(spec, pos, x, y, loc, scale, temperature)
        mom = mom - half * (-g)
    return pos, -mom, lt, g


class DATuner:
    """Dual-averaging step-size tuner, one independent state per chain; restatement of
    eeyore/tuners/hmcda_tuner.py:8-59 (python float64 arithmetic) for a given initial step e0."""

    def __init__(self, l, e0, n_chains, d=0.65, eub=None):
        self.l, self.d = float(l), float(d)
        self.m = math.log(10 * e0)                                   # set_m, :32-33
        self.logeub = None if eub is None else math.log(eub)
        self.barh = np.zeros(n_chains)
        self.logbare = np.zeros(n_chains)
        self.g, self.t0, self.k = 0.05, 10, 0.75
        self.step = np.full(n_chains, float(e0))
        self.num_steps = self.steps_for(self.step)

    def steps_for(self, e):
        return np.maximum(1, np.rint(self.l / e)).astype(np.int64)   # max(1, round(l / e)), :41-42 (round half to even)

    def tune(self, rate, idx, return_e=True):                        # :44-59
        it = idx + 1
        d_w = 1 / (it + self.t0)
        e_w = 1 / (it ** self.k)
        self.barh = (1 - d_w) * self.barh + d_w * (self.d - rate)
        loge = self.m - math.sqrt(it) * self.barh / self.g
        if self.logeub is not None:
            loge = np.minimum(loge, self.logeub)
        self.logbare = e_w * loge + (1 - e_w) * self.logbare
        self.step = np.exp(loge) if return_e else np.exp(self.logbare)
        self.num_steps = self.steps_for(self.step)


def hmc_run(spec: MLPSpec, x, y, loc, scale, theta0, z, u, step, num_steps, n_burnin=0,
            temperature=None, thin=1, tuner=None):
    """hmc.py:126-170.  z[t] is the momentum draw p0 of iteration t.  With a DATuner the step size and the number of
    leapfrog steps of every chain are adapted during burn-in (hmc.py:158-163)."""
    theta = np.array(np.atleast_2d(theta0), copy=True)
    dt = theta.dtype
    lt, g = log_target_grad(spec, theta, x, y, loc, scale, temperature)
    store = {}
    if tuner is not None:
        return _hmc_run_tuned(spec, x, y, loc, scale, theta, lt, g, z, u, n_burnin, temperature, thin, tuner)
    for t in range(z.shape[0]):
        p0 = z[t].astype(dt)
        h_cur = -lt + dt.type(0.5) * (p0 ** 2).sum(axis=1)                       # :137, :91-98
        prop, p1, lt_p, g_p = leapfrog(spec, x, y, loc, scale, theta, p0, step, num_steps, temperature)
        with np.errstate(invalid="ignore", over="ignore"):
            h_prop = -lt_p + dt.type(0.5) * (p1 ** 2).sum(axis=1)                # :141
            rate = np.exp(h_cur - h_prop)
            rate = np.where(np.isnan(rate), rate, np.minimum(rate, dt.type(1)))  # torch.min keeps NaN, :143-146
            acc = u[t].astype(dt) < rate                                         # :148
        theta = np.where(acc[:, None], prop, theta)
        g = np.where(acc[:, None], g_p, g)
        lt = np.where(acc, lt_p, lt)
        _collect(store, t, n_burnin, thin, sample=theta, target_val=lt, grad_val=g,
                 accepted=acc.astype(np.uint8))
    return _finish(store, dict(sample=theta, target_val=lt, grad_val=g))


def _hmc_run_tuned(spec, x, y, loc, scale, theta, lt, g, z, u, n_burnin, temperature, thin, tuner):
    """Per-chain step sizes: chains are advanced one by one (the leapfrog length differs between chains)."""
    dt = theta.dtype
    C = theta.shape[0]
    store = {}
    for t in range(z.shape[0]):
        acc = np.zeros(C, dtype=bool)
        rate = np.zeros(C)
        for c in range(C):
            p0 = z[t, c:c + 1].astype(dt)
            h_cur = -lt[c:c + 1] + dt.type(0.5) * (p0 ** 2).sum(axis=1)
            prop, p1, lt_p, g_p = leapfrog(spec, x, y, loc, scale, theta[c:c + 1], p0, float(tuner.step[c]),
                                           int(tuner.num_steps[c]), temperature)
            with np.errstate(invalid="ignore", over="ignore"):
                h_prop = -lt_p + dt.type(0.5) * (p1 ** 2).sum(axis=1)
                r = np.exp(h_cur - h_prop)
                r = np.where(np.isnan(r), r, np.minimum(r, dt.type(1)))
                a = u[t, c].astype(dt) < r[0]
            rate[c] = r[0]
            if a:
                theta[c], g[c], lt[c] = prop[0], g_p[0], lt_p[0]
            acc[c] = a
        if t < n_burnin:                                                          # hmc.py:158-163
            tuner.tune(rate, t, return_e=(t != n_burnin - 1))
        _collect(store, t, n_burnin, thin, sample=theta, target_val=lt, grad_val=g, accepted=acc.astype(np.uint8))
    return _finish(store, dict(sample=theta, target_val=lt, grad_val=g, step=tuner.step.copy(),
                               num_steps=tuner.num_steps.copy()))


# --------------------------------------------------------------------------------------
# SMMALA -- builder-defined algorithm, pinned by tests/golden/smmala_*.npz (make_golden.py:smmala_goldens)
# --------------------------------------------------------------------------------------

def fisher_metric(spec: MLPSpec, theta, x, y, loc, scale, temperature=None):
    """Expected Fisher information + prior precision for a binary MLP,
    G = T [ sum_i p_i (1-p_i) J_i J_i^T + diag(1/scale^2) ],  J_i = d g_L,i / d theta.
    Returns lt [C], grad [C,P], G [C,P,P]."""
    theta = np.atleast_2d(theta)
    dt = theta.dtype
    lt, g, J, p = log_target_grad(spec, theta, x, y, loc, scale, temperature, want_jac=True)
    w = p * (1 - p)
    G = np.einsum("cn,cni,cnj->cij", w, J, J)
    sc = np.asarray(scale, dtype=dt)
    G = G + np.diag(1 / (sc * sc))[None]
    if temperature is not None:
        G = dt.type(temperature) * G
    return lt, g, G


def _smmala_kernel(theta, g, G, step):
    """Cholesky G = R R^T (R lower); mean = theta + step/2 G^-1 grad; also returns ok flag."""
    C, P = theta.shape
    dt = theta.dtype
    R = np.zeros_like(G)
    ok = np.ones(C, dtype=bool)
    mean = np.full_like(theta, np.nan)
    for c in range(C):
        try:
            if not np.all(np.isfinite(G[c])):
                raise np.linalg.LinAlgError
            R[c] = np.linalg.cholesky(G[c])
            # solve G v = g via the two triangular systems
            v = np.linalg.solve(R[c].T, np.linalg.solve(R[c], g[c]))
            mean[c] = theta[c] + dt.type(0.5 * step) * v
        except np.linalg.LinAlgError:                # mirrors is_pos_def try/except, linalg/is_pos_def.py:5-9
            ok[c] = False
    return R, mean, ok


def _smmala_log_q(value, mean, R, step):
    """log N(value; mean, step G^-1) with G = R R^T:
    -(P/2) log(2 pi step) + sum log R_jj - |R^T (value-mean)|^2 / (2 step)."""
    P = value.shape[1]
    d = value - mean
    w = np.einsum("cji,cj->ci", R, d)               # R^T d
    logdet = np.log(np.einsum("cii->ci", R)).sum(axis=1)
    return -0.5 * P * math.log(2 * math.pi * step) + logdet - (w * w).sum(axis=1) / (2 * step)


def smmala_run(spec: MLPSpec, x, y, loc, scale, theta0, z, u, step, n_burnin=0, temperature=None, thin=1):
    """Simplified manifold MALA with the Fisher metric (MALA structure of mala.py:58-66 with a
    MultivariateNormalKernel(loc, scale_tril), kernels/multivariate_normal_kernel.py:11-19).
    Proposal theta' = mean(theta) + sqrt(step) R^-T z.  A Cholesky failure or non-finite value rejects."""
    theta = np.array(np.atleast_2d(theta0), copy=True)
    dt = theta.dtype
    C, P = theta.shape
    lt, g, G = fisher_metric(spec, theta, x, y, loc, scale, temperature)
    R, mean, ok = _smmala_kernel(theta, g, G, step)
    if not ok.all():
        raise RuntimeError("SMMALA: metric at the initial state is not positive definite")
    store = {}
    sq = dt.type(math.sqrt(step))
    for t in range(z.shape[0]):
        zz = z[t].astype(dt)
        prop = np.empty_like(theta)
        for c in range(C):
            prop[c] = mean[c] + sq * np.linalg.solve(R[c].T, zz[c])
        with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
            lt_p, g_p, G_p = fisher_metric(spec, prop, x, y, loc, scale, temperature)
            R_p, mean_p, ok_p = _smmala_kernel(prop, g_p, G_p, step)
            log_rate = lt_p - lt - _smmala_log_q(prop, mean, R, step) + _smmala_log_q(theta, mean_p, R_p, step)
            acc = ok_p & (np.log(u[t].astype(dt)) < log_rate)
        a1, a2 = acc[:, None], acc[:, None, None]
        theta = np.where(a1, prop, theta)
        g = np.where(a1, g_p, g)
        mean = np.where(a1, mean_p, mean)
        R = np.where(a2, R_p, R)
        lt = np.where(acc, lt_p, lt)
        _collect(store, t, n_burnin, thin, sample=theta, target_val=lt, grad_val=g,
                 accepted=acc.astype(np.uint8))
    return _finish(store, dict(sample=theta, target_val=lt, grad_val=g))


def am_run(spec: MLPSpec, x, y, loc, scale, theta0, z, u, n_burnin=0, cov0=None, l=0.05, b=1.0, c=1.0, t0=2,
           temperature=None, thin=1):
    """Adaptive Metropolis, eeyore/samplers/am.py:62-107 (offset = 0, no transform).  theta0 [C, P]; z [T, C, P];
    u [T, 2, C]: u[:, 0] is the mixture uniform (consumed only when idx + 1 > t0, am.py:69-70), u[:, 1] the accept uniform.
    A covariance estimate that is not positive definite raises (torch.linalg.cholesky does, am.py:74)."""
    theta = np.array(np.atleast_2d(theta0), copy=True)
    dt = theta.dtype
    nc, p = theta.shape
    cov0 = np.eye(p, dtype=dt) if cov0 is None else np.asarray(cov0, dtype=dt)
    cov = np.broadcast_to(cov0, (nc, p, p)).copy()
    mean = np.zeros((nc, p), dtype=dt)
    cov_sum = np.zeros((nc, p, p), dtype=dt)
    n_acc = np.zeros(nc, dtype=np.int64)
    lt = log_target(spec, theta, x, y, loc, scale, temperature)
    store = {}
    for idx in range(z.shape[0]):
        zt = z[idx].astype(dt)
        if idx + 1 > t0:                                                       # :69
            chol = np.linalg.cholesky(cov)                                     # raises LinAlgError like torch (:74)
            adapt = dt.type(b) * np.einsum("cij,cj->ci", chol, zt)
            prop = theta + np.where((u[idx, 0] < l)[:, None], dt.type(c) * zt, adapt)
        else:
            prop = theta + dt.type(c) * zt                                     # :76
        lt_p = log_target(spec, prop, x, y, loc, scale, temperature)
        with np.errstate(divide="ignore", invalid="ignore"):
            acc = np.log(u[idx, 1].astype(dt)) < lt_p - lt                     # :81
        theta = np.where(acc[:, None], prop, theta)
        lt = np.where(acc, lt_p, lt)
        if idx > 0:
            n_acc += acc                                                       # :87-88
        k = dt.type(idx + 1)
        mean = ((k - 1) * mean + theta) / k                                    # recursive_mean, :93-95
        cov_sum = cov_sum + theta[:, :, None] * theta[:, None, :]              # :96
        if idx + 1 >= t0:                                                      # :97-102
            kk = dt.type(idx)
            with np.errstate(divide="ignore", invalid="ignore"):
                rec = (cov_sum - (kk + 1) * (mean[:, :, None] * mean[:, None, :])) / kk
            cov = np.where((n_acc == 0)[:, None, None], cov0[None], rec)
        _collect(store, idx, n_burnin, thin, sample=theta, target_val=lt, accepted=acc.astype(np.uint8))
    return _finish(store, dict(sample=theta, target_val=lt, cov=cov, running_mean=mean, cov_sum=cov_sum, num_accepted=n_acc))


def ram_run(spec: MLPSpec, x, y, loc, scale, theta0, z, u, n_burnin=0, cov0=None, a=0.234, g=0.7, temperature=None, thin=1):
    """Robust adaptive Metropolis, eeyore/samplers/ram.py:39-70 (offset = 0).  theta0 [C, P]; z [T, C, P]; u [T, C]."""
    theta = np.array(np.atleast_2d(theta0), copy=True)
    dt = theta.dtype
    nc, p = theta.shape
    cov0 = np.eye(p, dtype=dt) if cov0 is None else np.asarray(cov0, dtype=dt)
    chol = np.broadcast_to(np.linalg.cholesky(cov0), (nc, p, p)).copy()        # :31-32
    lt = log_target(spec, theta, x, y, loc, scale, temperature)
    eye = np.eye(p, dtype=dt)
    store = {}
    for idx in range(z.shape[0]):
        zt = z[idx].astype(dt)
        prop = theta + np.einsum("cij,cj->ci", chol, zt)                       # :46
        lt_p = log_target(spec, prop, x, y, loc, scale, temperature)
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            log_rate = lt_p - lt
            acc = np.log(u[idx].astype(dt)) < log_rate                         # :51
            ex = np.exp(log_rate)
        alpha = np.where(ex < 1, ex, dt.type(1))                               # python min(1, nan) == 1
        theta = np.where(acc[:, None], prop, theta)
        lt = np.where(acc, lt_p, lt)
        h = min(1, p * (idx + 1) ** (-g))                                      # :61
        coef = (dt.type(h) * (alpha - dt.type(a)))[:, None, None]
        inner = eye[None] + coef * (zt[:, :, None] * zt[:, None, :]) / np.einsum("ci,ci->c", zt, zt)[:, None, None]
        chol = np.linalg.cholesky(chol @ inner @ np.transpose(chol, (0, 2, 1)))  # :62-66
        _collect(store, idx, n_burnin, thin, sample=theta, target_val=lt, accepted=acc.astype(np.uint8))
    return _finish(store, dict(sample=theta, target_val=lt, chol_cov=chol))
