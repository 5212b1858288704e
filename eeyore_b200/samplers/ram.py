"""Robust adaptive Metropolis; mirror of eeyore/samplers/ram.py:7-70.  Device code: eeyore_b200/csrc/adaptive.cuh (one warp
per chain; the Cholesky factor of the proposal is updated and re-factorised in shared memory every iteration)."""
import torch

from .. import _native as nv
from .am import _AdaptiveSampler


class RAM(_AdaptiveSampler):
    _entry = "eeyore_b200_ram_run"
    _kind = 1

    def __init__(self, model, theta0=None, dataloader=None, data0=None, counter=None, cov0=None, a=0.234, g=0.7, chain=None,
                 seed=None, thin=1):
        self.a, self.g = float(a), float(g)
        self._cov0_arg = cov0
        self.keys = ["sample", "target_val", "accepted"]
        self._init_native(model, theta0, dataloader, data0, counter, chain, seed, 0, thin)

    def _init_adaptive_state(self):
        """ram.py:31-36: chol_cov = cholesky(cov0)."""
        m = self.model
        p, c, dev = m.num_params(), self.num_chains, self._theta.device
        cov0 = torch.eye(p, dtype=m.dtype) if self._cov0_arg is None else torch.as_tensor(self._cov0_arg, dtype=m.dtype)
        if cov0.shape != (p, p):
            raise ValueError(f"cov0 must be [{p}, {p}]")
        self.cov0 = cov0
        chol = torch.linalg.cholesky(cov0.to(torch.float64)).to(m.dtype)          # once, on the host (set_cov, ram.py:31-32)
        self._adapt_state = chol.reshape(1, -1).repeat(c, 1).to(dev).contiguous()
        self._adapt_status = torch.zeros(c, dtype=torch.int32, device=dev)

    def _fill_params(self, p):
        p.adapt_p[0], p.adapt_p[1], p.adapt_p[2], p.adapt_t0 = self.a, self.g, 0.0, 0
        self._adapt_common(p)

    @property
    def chol_cov(self):
        p = self.model.num_params()
        low = torch.tril(self._adapt_state.reshape(-1, p, p))
        return low if self._batched else low[0]
