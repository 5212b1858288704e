"""Simplified manifold MALA with the expected-Fisher metric (builder-defined; the reference snapshot has no SMMALA,
SURVEY.md A.7).  Device code: eeyore_b200/csrc/smmala.cuh (metric accumulation, in-warp Cholesky, triangular solves)."""
from .native import NativeChainSampler


class SMMALA(NativeChainSampler):
    _entry = "eeyore_b200_smmala_run"

    def __init__(self, model, theta0=None, dataloader=None, data0=None, counter=None, step=0.1, chain=None, seed=None,
                 thin=1):
        self.step = step
        self.keys = ["sample", "target_val", "grad_val", "accepted"]
        self._init_native(model, theta0, dataloader, data0, counter, chain, seed, 0, thin)

    def _fill_params(self, p):
        p.step = float(self.step)

    def _spawn(self, theta0):
        return SMMALA(self.model, theta0=theta0, dataloader=self.dataloader, step=self.step, thin=self.thin)
