"""Mirror of eeyore/kernels/multivariate_normal_kernel.py:5-23 (the proposal shape of SMMALA)."""
import torch
from torch.distributions import MultivariateNormal


class MultivariateNormalKernel:
    def __init__(self, loc, scale_tril):
        self.set_density(loc, scale_tril)

    def set_density(self, loc, scale_tril):
        self.density = MultivariateNormal(loc, scale_tril=scale_tril)

    def set_density_params(self, loc, scale_tril=None):
        self.density = MultivariateNormal(loc, scale_tril=self.density.scale_tril if scale_tril is None else scale_tril)

    def log_prob(self, state):
        return torch.sum(self.density.log_prob(state))

    def sample(self):
        return self.density.sample()
