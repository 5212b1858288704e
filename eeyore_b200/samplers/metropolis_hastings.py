"""Random-walk Metropolis-Hastings; mirror of eeyore/samplers/metropolis_hastings.py:8-73.
Device code: eeyore_b200/csrc/samplers.cuh (mh_draw)."""
import torch

from ..kernels import NormalKernel
from .native import NativeChainSampler


class MetropolisHastings(NativeChainSampler):
    _entry = "eeyore_b200_mh_run"
    _uses_grad = False

    def __init__(self, model, theta0=None, dataloader=None, data0=None, counter=None, symmetric=True, kernel=None,
                 chain=None, scale=1.0, seed=None, lanes_per_chain=0, thin=1):
        self.symmetric = symmetric
        self.scale = float(scale)
        self.keys = ["sample", "target_val", "accepted"]
        self._init_native(model, theta0, dataloader, data0, counter, chain, seed, lanes_per_chain, thin)
        if kernel is not None:
            s = torch.as_tensor(kernel.density.scale).reshape(-1)
            if not torch.all(s == s[0]):
                raise ValueError("the native MH kernel uses one proposal scale for all parameters")
            self.scale = float(s[0])
        self.kernel = kernel or (self.default_kernel(self.current) if theta0 is not None else None)

    def default_kernel(self, state):
        """metropolis_hastings.py:25-28."""
        scale = torch.full([self.model.num_params()], self.scale, dtype=self.model.dtype, device=state["sample"].device)
        return NormalKernel(state["sample"], scale)

    def set_kernel(self, state, scale=None, scale_tril=None):
        self.kernel.set_density_params(state["sample"].clone().detach())

    def _fill_params(self, p):
        p.step, p.symmetric = self.scale, 1 if self.symmetric else 0

    def _publish_current(self):
        super()._publish_current()
        if getattr(self, "kernel", None) is not None:
            self.set_kernel(self.current)

    def _spawn(self, theta0):
        return MetropolisHastings(self.model, theta0=theta0, dataloader=self.dataloader, symmetric=self.symmetric,
                                  scale=self.scale, lanes_per_chain=self.lanes_per_chain, thin=self.thin)
