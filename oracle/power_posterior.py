"""TEST INFRASTRUCTURE (CPU oracle) -- numpy restatement of eeyore/samplers/power_posterior_sampler.py, batched over E
independent ensembles of K tempered chains.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it.

Pinned against the unmodified reference by tests/golden/pp_*.npz (oracle/make_golden.py: power_posterior_goldens), which
record the reference's proposal noise, its categorical neighbour draws and its accept uniforms in call order.

Follows, line by line:
  power_posterior_sampler.py:87-97   temperature ladder (default (i/K)^4) and per-sampler model temperature
  power_posterior_sampler.py:107-123 neighbour distribution  q_i(j) ~ exp(-b |j - i|), normalised by torch Categorical
  power_posterior_sampler.py:129-134 within-chain moves: one MH / MALA draw per level (metropolis_hastings.py:41-73, mala.py:46-82)
  power_posterior_sampler.py:136-142 between-chain log-rate
  power_posterior_sampler.py:144-167 swap (reset = re-evaluate at the level's own temperature) or revert
  power_posterior_sampler.py:173-182 draw: within moves; between moves when counter.idx % between_step == 0; then save
"""
import numpy as np

from .mlp import MLPSpec, log_target, log_target_grad
from .samplers import _normal_log_prob_sum


def default_temperatures(k):
    """power_posterior_sampler.py:91-92."""
    return [(i / k) ** 4 for i in range(1, k + 1)]


def categorical_log_probs(k, b):
    """LQ[i, j] = log-probability of drawing level j from level i's neighbour distribution (diagonal = -inf).
    power_posterior_sampler.py:107-121 + torch.distributions.Categorical(probs): normalise, clamp to [eps, 1 - eps], log."""
    eb = np.exp(-b)
    lq = np.full((k, k), -np.inf)
    eps = np.finfo(np.float64).eps
    for i in range(k):
        js = [j for j in range(k) if j != i]
        den = eb * (2 - eb ** i - eb ** (k - 1 - i)) / (1 - eb)
        p = np.array([eb ** abs(j - i) / den for j in js], dtype=np.float64)
        p = p / p.sum()
        lq[i, js] = np.log(np.clip(p, eps, 1 - eps))
    return lq


def _mh_step(spec, x, y, loc, scale, theta, lt, z, u, temp, prop_scale):
    """metropolis_hastings.py:41-73 (symmetric kernel) from a carried (theta, target) state."""
    prop = theta + prop_scale * z
    lt_p = log_target(spec, prop, x, y, loc, scale, temp)
    with np.errstate(divide="ignore", invalid="ignore"):
        acc = np.log(u) < lt_p - lt
    return np.where(acc[:, None], prop, theta), np.where(acc, lt_p, lt)


def _mala_step(spec, x, y, loc, scale, theta, lt, g, z, u, step, temp):
    """mala.py:46-82 from a carried (theta, target, gradient) state."""
    dt = theta.dtype
    half_step = dt.type(0.5 * step)
    ps = np.full(theta.shape[1], np.sqrt(step), dtype=dt)
    mean_c = theta + half_step * g
    prop = mean_c + ps * z
    lt_p, g_p = log_target_grad(spec, prop, x, y, loc, scale, temp)
    with np.errstate(invalid="ignore", divide="ignore"):
        log_rate = lt_p - lt
        log_rate = log_rate - _normal_log_prob_sum(prop, mean_c, ps)
        mean_p = prop + half_step * g_p
        log_rate = log_rate + _normal_log_prob_sum(theta, mean_p, ps)
        acc = np.log(u) < log_rate
    return np.where(acc[:, None], prop, theta), np.where(acc, lt_p, lt), np.where(acc[:, None], g_p, g)


def power_posterior_run(spec: MLPSpec, x, y, loc, scale, theta0, kinds, kwargs, z, u, j_tape, u_between, temperatures=None,
                        between_step=10, b=0.5, n_burnin=0):
    """theta0 [E, P] (every level starts there, power_posterior_sampler.py:69-84); kinds[k] in {"mh", "mala"};
    kwargs[k]: {"step": ...} for MALA, {"prop_scale": ...} for MH; z [T, K, E, P]; u [T, K, E];
    j_tape [NB, K, E] neighbour draws and u_between [NB, K, E] accept uniforms of the NB between-chain sweeps.
    Returns per-level saved samples [K, n_saved, E, P], targets [K, n_saved, E], the swap decisions [NB, K, E] and the final
    states."""
    theta0 = np.array(np.atleast_2d(theta0), dtype=np.float64)
    e, p = theta0.shape
    k = len(kinds)
    temps = list(temperatures) if temperatures is not None else default_temperatures(k)
    lq = categorical_log_probs(k, b)
    theta = np.stack([theta0.copy() for _ in range(k)])                        # [K, E, P]
    # Reference quirk kept on purpose: init_samplers (:33) evaluates every sampler's initial target / gradient BEFORE
    # set_temperature (:35) assigns the ladder, i.e. with temperature None; the stale untempered values are used until the
    # level's first accepted move or swap.
    lt0, g0 = log_target_grad(spec, theta0, x, y, loc, scale, None)
    lt = np.stack([lt0.copy() for _ in range(k)])
    g = np.stack([g0.copy() for _ in range(k)])
    ar = np.arange(e)
    samples, targets, swaps = [], [], []
    nb = 0
    for t in range(z.shape[0]):
        for m in range(k):                                                     # within-chain moves (:132-134)
            if kinds[m] == "mh":
                theta[m], lt[m] = _mh_step(spec, x, y, loc, scale, theta[m], lt[m], z[t, m], u[t, m], temps[m],
                                           kwargs[m].get("prop_scale", 1.0))
            else:
                theta[m], lt[m], g[m] = _mala_step(spec, x, y, loc, scale, theta[m], lt[m], g[m], z[t, m], u[t, m],
                                                   kwargs[m]["step"], temps[m])
        if t % between_step == 0:                                              # :176
            sw = np.zeros((k, e), dtype=np.uint8)
            for i in range(k):                                                 # :169-171, sequential in i
                j = j_tape[nb, i].astype(np.int64)
                th_i, th_j = theta[i].copy(), theta[j, ar].copy()
                lt_i, lt_j = lt[i].copy(), lt[j, ar].copy()
                t_j = np.array([temps[jj] for jj in j])
                g_i = g[i].copy()
                cross_i, gx_i = log_target_grad(spec, th_j, x, y, loc, scale, temps[i])   # sampler_i.model at theta_j
                raw_ll, raw_lp = _raw_parts(spec, th_i, x, y, loc, scale)
                cross_j = t_j * raw_ll + t_j * raw_lp                          # sampler_j.model.log_target(theta_i)
                _, graw_i = log_target_grad(spec, th_i, x, y, loc, scale, None)
                gx_j = t_j[:, None] * graw_i                                   # gradient of level j's target at theta_i
                with np.errstate(invalid="ignore", divide="ignore"):
                    log_rate = lq[j, i] - lq[i, j] - lt_i - lt_j + cross_i + cross_j        # :137-142
                    acc = np.log(u_between[nb, i]) < log_rate                  # :162
                theta[i] = np.where(acc[:, None], th_j, th_i)                  # swap_states: reset re-evaluates target
                lt[i] = np.where(acc, cross_i, lt_i)                           # and gradient at the level's temperature
                g[i] = np.where(acc[:, None], gx_i, g_i)
                for m in range(k):
                    hit = acc & (j == m)
                    theta[m] = np.where(hit[:, None], th_i, theta[m])
                    lt[m] = np.where(hit, cross_j, lt[m])
                    g[m] = np.where(hit[:, None], gx_j, g[m])
                sw[i] = acc
            swaps.append(sw)
            nb += 1
        if t >= n_burnin:                                                      # :179-181: saved after the between moves
            samples.append(theta.copy())
            targets.append(lt.copy())
    out = dict(final_sample=theta, final_target=lt, swaps=np.stack(swaps) if swaps else np.zeros((0, k, e), np.uint8))
    if samples:
        out["sample"] = np.stack(samples, axis=1)                              # [K, n_saved, E, P]
        out["target_val"] = np.stack(targets, axis=1)
    return out


def _raw_parts(spec, theta, x, y, loc, scale):
    """log-likelihood and log-prior without temperature (bayesian_model.py:30-50), separately."""
    from .mlp import log_lik, log_prior
    return log_lik(spec, theta, x, y), log_prior(theta, loc, scale)
