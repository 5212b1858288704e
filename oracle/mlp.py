"""Oracle (test infrastructure): closed-form MLP log-target and gradient, batched over chains.

numpy restatement of
  eeyore/models/model.py:44-55          (flat theta layout: per layer W row-major [out,in], then b)
  eeyore/models/mlp.py:45-50            (forward: fc -> optional sigmoid)
  eeyore/stats/loss.py:1-11             (naive binary cross-entropy on probabilities, reduction='sum')
  eeyore/constants/constants.py:15-18   (binary / multiclass loss table)
  eeyore/models/bayesian_model.py:30-56 (log_lik, log_prior, log_target, temperature on both terms)
  eeyore/models/log_target_model.py:15-23 (gradient of log_target over theta; here in closed form)

All functions take ``theta`` of shape [C, P] (C independent chains) and return
per-chain values; C = 1 reproduces the reference's single evaluation.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

SIGMOID = "sigmoid"
BINARY = "binary_classification"
MULTICLASS = "multiclass_classification"


@dataclass
class MLPSpec:
    """Architecture description; mirrors mlp.Hyperparameters (eeyore/models/mlp.py:9-19)."""

    dims: Sequence[int]
    loss: str = BINARY
    bias: Optional[Sequence[bool]] = None
    activations: Optional[Sequence[Optional[str]]] = None
    offsets: List[int] = field(default_factory=list, init=False)

    def __post_init__(self):
        nl = len(self.dims) - 1
        if self.bias is None:
            self.bias = [True] * nl
        if self.activations is None:
            last = SIGMOID if self.loss == BINARY else None
            self.activations = [SIGMOID] * (nl - 1) + [last]
        if len(self.dims) < 3 or len(self.dims) != len(self.activations) + 1:
            raise ValueError  # mlp.py:15-19
        # layer start offsets, eeyore/models/mlp.py:72-79
        s = 0
        self.offsets = []
        for l in range(nl):
            self.offsets.append(s)
            s += (self.dims[l] + (1 if self.bias[l] else 0)) * self.dims[l + 1]
        self.num_params = s

    @property
    def num_layers(self):
        return len(self.dims) - 1


def _sigmoid(g):
    with np.errstate(over="ignore"):
        return 1.0 / (1.0 + np.exp(-g))


def unpack(spec: MLPSpec, theta: np.ndarray):
    """theta [C,P] -> list of (W [C,dout,din], b [C,dout] | None).  model.py:44-55."""
    out = []
    for l in range(spec.num_layers):
        din, dout = spec.dims[l], spec.dims[l + 1]
        s = spec.offsets[l]
        W = theta[:, s:s + din * dout].reshape(-1, dout, din)
        b = theta[:, s + din * dout:s + din * dout + dout] if spec.bias[l] else None
        out.append((W, b))
    return out


def forward(spec: MLPSpec, theta: np.ndarray, x: np.ndarray):
    """Returns list h[0..L] with h[l] of shape [C,N,d_l]; h[L] is the network output
    (probability for a sigmoid head, logits for a None head).  mlp.py:45-50."""
    theta = np.atleast_2d(theta)
    C = theta.shape[0]
    h = [np.broadcast_to(x[None], (C,) + x.shape)]
    for l, (W, b) in enumerate(unpack(spec, theta)):
        g = np.einsum("cni,coi->cno", h[-1], W)
        if b is not None:
            g = g + b[:, None, :]
        h.append(_sigmoid(g) if spec.activations[l] == SIGMOID else g)
    return h


def _loss_seed(spec: MLPSpec, out: np.ndarray, y: np.ndarray):
    """Per-chain log-likelihood ll [C] and d ll / d g_L [C,N,d_L] (pre-activation of the last layer).

    binary    : stats/loss.py:2 on probabilities, including its 0 * (-inf) = NaN behaviour
                (SURVEY.md A.8); seed follows what autograd produces from that form:
                (y/p - (1-y)/(1-p)) * (1-p) * p.
    multiclass: constants.py:17, CrossEntropyLoss(sum) against argmax(y, 1).
    """
    dt = out.dtype
    if spec.loss == BINARY:
        p = out
        yy = y.reshape(1, -1, 1).astype(dt)
        one = dt.type(1)
        with np.errstate(divide="ignore", invalid="ignore"):
            ll = (np.log(p) * yy + np.log(one - p) * (one - yy)).sum(axis=(1, 2))
            seed = (yy * (one / p) - (one - yy) * (one / (one - p))) * (one - p) * p
        return ll, seed
    if spec.loss == MULTICLASS:
        g = out
        c = np.argmax(y, axis=1)
        m = g.max(axis=2, keepdims=True)
        e = np.exp(g - m)
        s = e.sum(axis=2, keepdims=True)
        logsm = g - m - np.log(s)
        n_idx = np.arange(g.shape[1])
        ll = logsm[:, n_idx, c].sum(axis=1)
        seed = -(e / s)
        seed[:, n_idx, c] += 1
        return ll, seed
    raise ValueError(spec.loss)


def log_lik(spec: MLPSpec, theta, x, y, temperature=None):
    """bayesian_model.py:30-35."""
    theta = np.atleast_2d(theta)
    h = forward(spec, theta, x)
    ll, _ = _loss_seed(spec, h[-1], y)
    return ll if temperature is None else theta.dtype.type(temperature) * ll


def log_prior(theta, loc, scale, temperature=None, want_grad=False):
    """Vector Normal prior, bayesian_model.py:46-50 with torch.distributions.Normal.log_prob:
    -(v-loc)^2/(2 scale^2) - log(scale) - log(sqrt(2 pi)), summed over parameters."""
    theta = np.atleast_2d(theta)
    dt = theta.dtype
    loc = np.asarray(loc, dtype=dt)
    scale = np.asarray(scale, dtype=dt)
    var = scale * scale
    d = theta - loc
    lp = (-(d * d) / (2 * var) - np.log(scale) - dt.type(math.log(math.sqrt(2 * math.pi)))).sum(axis=1)
    g = -d / var
    if temperature is not None:
        lp = dt.type(temperature) * lp
        g = dt.type(temperature) * g
    return (lp, g) if want_grad else lp


def log_target_grad(spec: MLPSpec, theta, x, y, loc, scale, temperature=None, want_jac=False):
    """log_target [C] and its gradient [C,P]; log_target_model.py:20-23 in closed form
    (SURVEY.md A.3-A.5).  With ``want_jac`` also returns the per-row Jacobian of the
    last pre-activation, J [C,N,P] (binary nets only; used by the SMMALA metric)."""
    theta = np.atleast_2d(theta)
    dt = theta.dtype
    x = np.asarray(x, dtype=dt)
    C = theta.shape[0]
    layers = unpack(spec, theta)
    h = forward(spec, theta, x)
    ll, seed = _loss_seed(spec, h[-1], y)

    def backprop(delta):
        """delta [C,N,d_L] on the last pre-activation -> per-row parameter cotangents [C,N,P]."""
        per_row = np.zeros((C, x.shape[0], spec.num_params), dtype=dt)
        for l in range(spec.num_layers - 1, -1, -1):
            W, b = layers[l]
            din, dout = spec.dims[l], spec.dims[l + 1]
            s = spec.offsets[l]
            per_row[:, :, s:s + din * dout] = (delta[:, :, :, None] * h[l][:, :, None, :]).reshape(C, -1, din * dout)
            if b is not None:
                per_row[:, :, s + din * dout:s + din * dout + dout] = delta
            if l > 0:
                back = np.einsum("cno,coi->cni", delta, W)
                if spec.activations[l - 1] == SIGMOID:
                    back = back * (1 - h[l]) * h[l]
                delta = back
        return per_row

    with np.errstate(invalid="ignore"):
        gll = backprop(seed).sum(axis=1)
    lp, glp = log_prior(theta, loc, scale, want_grad=True)
    lt = ll + lp
    g = gll + glp
    if temperature is not None:
        # both terms are scaled, bayesian_model.py:33-34,48-49
        t = dt.type(temperature)
        lt = t * ll + t * lp
        g = t * gll + t * glp
    if want_jac:
        assert spec.loss == BINARY and spec.dims[-1] == 1
        J = backprop(np.ones_like(seed))
        return lt, g, J, h[-1][:, :, 0]
    return lt, g


def log_target(spec: MLPSpec, theta, x, y, loc, scale, temperature=None):
    """bayesian_model.py:52-56."""
    theta = np.atleast_2d(theta)
    ll = log_lik(spec, theta, np.asarray(x, dtype=theta.dtype), y, temperature)
    lp = log_prior(theta, loc, scale, temperature)
    return ll + lp
