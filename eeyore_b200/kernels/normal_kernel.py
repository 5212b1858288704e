"""Proposal-density handle; mirror of eeyore/kernels/normal_kernel.py:5-23 + normalized_kernel.py:14-19.

The fused samplers draw proposals and evaluate q(.|.) on the device (eeyore_b200/csrc/samplers.cuh); this object
only carries (loc, scale) so `sampler.kernel` keeps the reference's attributes.
"""
import torch
from torch.distributions import Normal


class NormalKernel:
    def __init__(self, loc, scale):
        self.set_density(loc, scale)

    def set_density(self, loc, scale):
        self.density = Normal(loc, scale)

    def set_density_params(self, loc, scale=None):
        self.density.loc = loc
        if scale is not None:
            self.density.scale = scale

    def log_prob(self, state):
        return torch.sum(self.density.log_prob(state))

    def sample(self):
        return self.density.sample()

    def k(self, x1, x2, scale=None):
        self.set_density_params(x2, scale=scale)
        return self.log_prob(x1).exp()
