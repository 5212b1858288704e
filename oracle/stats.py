"""Oracle (test infrastructure): sample covariance, initial-sequence Monte Carlo covariance, multivariate ESS, ACF.

numpy restatement of
  eeyore/stats/cov.py:5-15           (sample covariance, n-1 denominator)
  eeyore/linalg/is_pos_def.py:3-11   (exact symmetry and Cholesky success)
  eeyore/stats/inse_mc_cov.py:9-83   (INSE estimator, adjust=False path)
  eeyore/stats/multi_ess.py:6-14     (n (det cov / det inse)^(1/p))
The lag sums are evaluated as matrix products instead of the reference's python loop of torch.ger
outer products; results agree to rounding.

ACF is **builder-defined (parity unpinned)**: it lives in the un-vendored ``kanga`` package, absent from the
reference snapshot; the definition is SURVEY.md A.10.
"""
from __future__ import annotations

import numpy as np


def cov(x):
    """x [n,p] -> [p,p]; cov.py:13-15."""
    xc = x - x.mean(axis=0, keepdims=True)
    return xc.T @ xc / (x.shape[0] - 1)


def is_pos_def(m):
    """is_pos_def.py:3-11."""
    if not np.array_equal(m, m.T):
        return False
    if not np.all(np.isfinite(m)):
        return False
    try:
        np.linalg.cholesky(m)
        return True
    except np.linalg.LinAlgError:
        return False


def _gamma_pair(xc, m):
    n = xc.shape[0]
    g0 = xc[: n - 2 * m].T @ xc[2 * m:] / n               # inse_mc_cov.py:24-26
    g1 = xc[: n - 2 * m - 1].T @ xc[2 * m + 1:] / n       # :28-30
    gam = g0 + g1
    return g0, (gam + gam.T) / 2                          # :32-33


def inse_mc_cov(x, return_info=False):
    """inse_mc_cov.py:9-83 with adjust=False.  Raises RuntimeError('Not enough samples') like :44-45."""
    xc = x - x.mean(axis=0)
    n, p = x.shape
    ub = n // 2
    sn = ub
    sig = None
    for m in range(ub):
        g0, gam = _gamma_pair(xc, m)
        sig = -g0 + 2 * gam if m == 0 else sig + 2 * gam  # :35-38
        if is_pos_def(sig):
            sn = m
            break
    if sn > ub - 1:
        raise RuntimeError("Not enough samples")
    last = np.linalg.det(sig)
    m_last = sn
    for m in range(sn + 1, ub):
        _, gam = _gamma_pair(xc, m)
        sig1 = sig + 2 * gam
        cur = np.linalg.det(sig1)
        if cur <= last:                                   # :66-69
            break
        sig, last, m_last = sig1, cur, m
    if return_info:
        return sig, dict(sn=sn, m_last=m_last)
    return sig


def multi_ess(x):
    """multi_ess.py:6-14."""
    n, p = x.shape
    return n * (np.linalg.det(cov(x)) / np.linalg.det(inse_mc_cov(x))) ** (1.0 / p)


def acf(x, max_lag):
    """Builder-defined: rho_k[j] = sum_t (x_t - xbar)(x_{t+k} - xbar) / sum_t (x_t - xbar)^2,
    per parameter, k = 0..max_lag.  x [n,p] -> [max_lag+1, p]."""
    xc = x - x.mean(axis=0)
    n = x.shape[0]
    den = (xc * xc).sum(axis=0)
    out = np.empty((max_lag + 1, x.shape[1]), dtype=x.dtype)
    for k in range(max_lag + 1):
        out[k] = (xc[: n - k] * xc[k:]).sum(axis=0) / den
    return out
