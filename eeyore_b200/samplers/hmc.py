"""Hamiltonian Monte Carlo; mirror of eeyore/samplers/hmc.py:8-170.
Device code: eeyore_b200/csrc/samplers.cuh (hmc_draw): the whole leapfrog trajectory and the accept test run inside
the kernel.  HMCDATuner (hmc.py:17-28,158-163) is a SURVEY.md section 8(f) "next" row and is not implemented."""
from .native import NativeChainSampler


class HMC(NativeChainSampler):
    _entry = "eeyore_b200_hmc_run"

    def __init__(self, model, theta0=None, dataloader=None, data0=None, counter=None, step=0.1, num_steps=10,
                 tuner=None, chain=None, seed=None, lanes_per_chain=0, thin=1):
        if tuner is not None:
            raise NotImplementedError("HMCDATuner is not part of the native hot path yet (SURVEY.md 8(f) row 1)")
        self.tuner = None
        self.step, self.num_steps = step, num_steps
        self.keys = ["sample", "target_val", "grad_val", "momentum", "hamiltonian", "accepted"]
        self._init_native(model, theta0, dataloader, data0, counter, chain, seed, lanes_per_chain, thin)

    def _fill_params(self, p):
        p.step, p.num_steps = float(self.step), int(self.num_steps)

    def evals_per_iteration(self):
        """Gradient evaluations the kernel executes per iteration (the reference executes num_steps + 1, hmc.py:104-118;
        its first one recomputes the cached current gradient)."""
        return int(self.num_steps)

    def _spawn(self, theta0):
        return HMC(self.model, theta0=theta0, dataloader=self.dataloader, step=self.step, num_steps=self.num_steps,
                   lanes_per_chain=self.lanes_per_chain, thin=self.thin)
