"""Adaptive Metropolis; mirror of eeyore/samplers/am.py:8-107.  Device code: eeyore_b200/csrc/adaptive.cuh (one warp per
chain; the covariance estimate, its running sums and its Cholesky factor live in shared memory for the whole run).

Deviations: `transform` (a Python callable applied to the covariance every iteration, e.g. softabs) cannot run inside the
fused kernel and raises ValueError; a covariance estimate that is not positive definite raises RuntimeError after the run
(the reference's torch.linalg.cholesky raises at that iteration, am.py:74) -- as there, choose t0 large enough for the plain
estimator.  The accept-uniform tape holds two values per iteration (mixture test am.py:70, accept test am.py:81)."""
import ctypes as C

import torch

from .. import _native as nv
from .native import NativeChainSampler


class _AdaptiveSampler(NativeChainSampler):
    _uses_grad = False
    _kind = None

    def _init_adaptive_state(self):
        raise NotImplementedError

    def set_current(self, theta, data=None):
        out = super().set_current(theta, data=data)
        self._init_adaptive_state()
        return out

    def _adapt_common(self, p):
        p.adapt_iter0 = int(self.counter.idx)
        p.adapt_state = self._adapt_state.data_ptr()
        p.adapt_status = self._adapt_status.data_ptr()

    def _run_fused(self, n_iters):
        super()._run_fused(n_iters)
        self.check_status()

    def check_status(self):
        """torch.linalg.cholesky raises inside the reference's draw when the proposal covariance is not positive definite."""
        bad = self._adapt_status.nonzero()
        if bad.numel():
            c = int(bad[0, 0].item())
            raise RuntimeError(f"linalg.cholesky: the proposal covariance of chain {c} is not positive-definite "
                               f"(iteration {int(self._adapt_status[c].item()) - 1})")

    def set_noise_tape(self, z, u):
        """z [T, P] or [T, C, P]; u [T, U] or [T, U, C] with U uniforms per iteration (AM: 2, RAM: 1)."""
        m = self.model
        z, u = m._to_dev(z), m._to_dev(u)
        self._tape = [z.reshape(z.shape[0], -1, m.num_params()), u.reshape(u.shape[0], -1), 0]


class AM(_AdaptiveSampler):
    _entry = "eeyore_b200_am_run"
    _kind = 0

    def __init__(self, model, theta0=None, dataloader=None, data0=None, counter=None, cov0=None, l=0.05, b=1., c=1., t0=2,
                 transform=None, chain=None, seed=None, thin=1):
        if transform is not None:
            raise ValueError("the device AM sampler cannot apply a Python transform to the covariance inside the fused run")
        self.l, self.b, self.c, self.t0 = float(l), float(b), float(c), int(t0)
        self._cov0_arg = cov0
        self.keys = ["sample", "target_val", "accepted"]
        self._init_native(model, theta0, dataloader, data0, counter, chain, seed, 0, thin)

    def _init_adaptive_state(self):
        """am.py:40-52: cov = cov0, running mean and cov_sum zero, num_accepted = 0."""
        m = self.model
        p, c, dev = m.num_params(), self.num_chains, self._theta.device
        cov0 = torch.eye(p, dtype=m.dtype) if self._cov0_arg is None else torch.as_tensor(self._cov0_arg, dtype=m.dtype)
        if cov0.shape != (p, p):
            raise ValueError(f"cov0 must be [{p}, {p}]")
        self.cov0 = cov0.to(dev).contiguous()
        n = int(nv.lib().eeyore_b200_adapt_state_len(m.handle(), self._kind))
        st = torch.zeros(c, n, dtype=m.dtype, device=dev)
        st[:, p + p * p:p + 2 * p * p] = self.cov0.reshape(-1)
        self._adapt_state = st
        self._adapt_status = torch.zeros(c, dtype=torch.int32, device=dev)

    def reset(self, theta, data=None, reset_counter=True, reset_chain=True):
        super().reset(theta, data=data, reset_counter=reset_counter, reset_chain=reset_chain)

    def _fill_params(self, p):
        p.adapt_p[0], p.adapt_p[1], p.adapt_p[2], p.adapt_t0 = self.l, self.b, self.c, self.t0
        p.adapt_cov0 = self.cov0.data_ptr()
        self._adapt_common(p)

    @property
    def cov(self):
        """Current covariance estimate(s), [P, P] or [C, P, P] (lower triangle mirrored)."""
        p = self.model.num_params()
        low = torch.tril(self._adapt_state[:, p + p * p:p + 2 * p * p].reshape(-1, p, p))
        full = low + torch.tril(low, -1).transpose(1, 2)
        return full if self._batched else full[0]

    @property
    def num_accepted(self):
        v = self._adapt_state[:, -1].to(torch.int64)
        return v if self._batched else int(v[0].item())
