"""Metropolis-adjusted Langevin algorithm; mirror of eeyore/samplers/mala.py:9-82.
Device code: eeyore_b200/csrc/samplers.cuh (mala_draw)."""
import numpy as np
import torch

from ..kernels import NormalKernel
from .native import NativeChainSampler


class MALA(NativeChainSampler):
    _entry = "eeyore_b200_mala_run"

    def __init__(self, model, theta0=None, dataloader=None, data0=None, counter=None, step=0.1, kernel=None,
                 chain=None, seed=None, lanes_per_chain=0, thin=1):
        self.step = step
        self.keys = ["sample", "target_val", "grad_val", "accepted"]
        self._init_native(model, theta0, dataloader, data0, counter, chain, seed, lanes_per_chain, thin)
        if kernel is not None:
            raise ValueError("the native MALA builds its own N(theta + step/2 grad, step I) proposal (mala.py:35-41)")
        self.kernel = self.default_kernel(self.current) if theta0 is not None else None

    def kernel_mean(self, state):
        """mala.py:35-36."""
        return state["sample"] + 0.5 * self.step * state["grad_val"]

    def default_kernel(self, state):
        """mala.py:38-41."""
        scale = torch.full([self.model.num_params()], np.sqrt(self.step), dtype=self.model.dtype,
                           device=state["sample"].device)
        return NormalKernel(self.kernel_mean(state), scale)

    def set_kernel(self, state):
        self.kernel.set_density_params(self.kernel_mean(state))

    def _fill_params(self, p):
        p.step = float(self.step)

    def _publish_current(self):
        super()._publish_current()
        if getattr(self, "kernel", None) is not None:
            self.set_kernel(self.current)

    def _spawn(self, theta0):
        return MALA(self.model, theta0=theta0, dataloader=self.dataloader, step=self.step,
                    lanes_per_chain=self.lanes_per_chain, thin=self.thin)
