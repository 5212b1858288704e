"""Hamiltonian Monte Carlo; mirror of eeyore/samplers/hmc.py:8-170.
Device code: eeyore_b200/csrc/samplers.cuh (hmc_draw, da_tune): the whole leapfrog trajectory, the accept test and --
with an HMCDATuner -- the dual-averaging step-size adaptation of the burn-in run inside the kernel, per chain."""
import math

import torch

from .native import NativeChainSampler


class HMC(NativeChainSampler):
    _entry = "eeyore_b200_hmc_run"

    def __init__(self, model, theta0=None, dataloader=None, data0=None, counter=None, step=0.1, num_steps=10,
                 tuner=None, chain=None, seed=None, lanes_per_chain=0, thin=1):
        self.tuner = tuner
        self._tuner_state = None
        self.step, self.num_steps = step, num_steps
        self._step0, self._num_steps0 = None, None    # scalar values the per-chain tuner state starts from
        self.keys = ["sample", "target_val", "grad_val", "momentum", "hamiltonian", "accepted"]
        self._init_native(model, theta0, dataloader, data0, counter, chain, seed, lanes_per_chain, thin)
        if tuner is not None:                                    # hmc.py:17-28
            if tuner.e0 is None:
                if theta0 is None or self.num_chains != 1:
                    raise ValueError("HMCDATuner without e0 needs a single initial state (init_step heuristic, hmc.py:38-77)")
                self.init_step()
                if tuner.eub is not None:
                    self.step = min(tuner.eub, self.step)
                tuner.set_m(self.step)
            else:
                self.step = tuner.e0
            self.num_steps = tuner.num_steps(self.step)

    # -- step-size heuristic of the reference (setup time only; a handful of native evaluations) --------------------
    def init_step(self):
        """hmc.py:38-77 (NUTS-paper heuristic): double / halve the step until the one-step acceptance crosses 1/2."""
        m = self.model
        xd, yd = self._data_dev
        theta = self._theta[0].contiguous()
        mom = torch.randn(m.num_params(), dtype=m.dtype, device=theta.device)

        def one_step(step):
            lt0, g0 = m._eval(theta[None], xd, yd)
            p = mom + 0.5 * step * g0[0]
            lt1, g1 = m._eval((theta + step * p)[None], xd, yd)
            p = p + 0.5 * step * g1[0]
            h0 = -lt0[0] + 0.5 * torch.sum(mom ** 2)
            h1 = -lt1[0] + 0.5 * torch.sum(p ** 2)
            return torch.exp(h0 - h1).item()

        self.step, self.num_steps = 1.0, 1
        ratio = one_step(self.step)
        a = 1 if ratio > 0.5 else -1
        guard = 0
        while ratio ** a > 2.0 ** (-a) and guard < 200:
            self.step = (2.0 ** a) * self.step
            ratio = one_step(self.step)
            guard += 1

    def _scalar_step(self):
        """(step, num_steps) as python scalars.  After a batched tuned run self.step / self.num_steps are per-chain tensors;
        the scalars the tuner state was started from are kept aside for that case."""
        if isinstance(self.step, torch.Tensor):
            return self._step0, self._num_steps0
        return float(self.step), int(self.num_steps)

    def _on_new_state(self):
        """set_current / reset.  One chain: the tuner state carries over, as the reference's HMCDATuner object does.
        C chains: the per-chain dual-averaging state ([4, C]: barh, logbare, step, num_steps) belongs to the previous set
        of chains -- drop it (a buffer of another C would be indexed out of bounds by the kernel) and go back to the scalar
        step the tuner was started from."""
        st = getattr(self, "_tuner_state", None)
        if st is not None and (self._batched or st.shape[1] != self.num_chains):
            if isinstance(self.step, torch.Tensor):
                self.step, self.num_steps = self._step0, self._num_steps0
            self._tuner_state = None

    def _tuner_buffers(self):
        c = self.num_chains
        if self._tuner_state is None or self._tuner_state.shape != (4, c):
            self._step0, self._num_steps0 = self._scalar_step()
            st = torch.zeros(4, c, dtype=torch.float64, device=self._theta.device)
            st[2] = self._step0
            st[3] = float(self._num_steps0)
            self._tuner_state = st
        return self._tuner_state

    def _fill_params(self, p):
        p.step, p.num_steps = self._scalar_step()
        if self.tuner is not None:
            t = self.tuner
            st = self._tuner_buffers()
            p.tuner_l, p.tuner_d, p.tuner_m = float(t.l), float(t.d), float(t.m)
            p.tuner_has_eub = 0 if t.eub is None else 1
            p.tuner_logeub = 0.0 if t.eub is None else math.log(t.eub)
            nb = self.counter.num_burnin_iters or 0
            p.tuner_iter0 = self.counter.idx
            p.tuner_burnin = max(0, nb - self.counter.idx)
            p.tuner_state = st.data_ptr()

    def _publish_current(self):
        super()._publish_current()
        if getattr(self, "tuner", None) is not None and self._tuner_state is not None:
            st = self._tuner_state
            if self._batched:
                self.step, self.num_steps = st[2].clone(), st[3].to(torch.int64)
            else:
                self.step, self.num_steps = st[2, 0].item(), int(st[3, 0].item())
                self.tuner.barh, self.tuner.logbare = st[0, 0].item(), st[1, 0].item()

    def evals_per_iteration(self):
        """Gradient evaluations the kernel executes per iteration (the reference executes num_steps + 1, hmc.py:104-118;
        its first one recomputes the cached current gradient)."""
        return self.num_steps

    def _spawn(self, theta0):
        from ..tuners import HMCDATuner
        t = None
        if self.tuner is not None:
            t = HMCDATuner(self.tuner.l, e0=self.tuner.e0 if self.tuner.e0 is not None else self._scalar_step()[0],
                           d=self.tuner.d, eub=self.tuner.eub)
        step, num_steps = self._scalar_step()
        return HMC(self.model, theta0=theta0, dataloader=self.dataloader, step=step, num_steps=num_steps,
                   tuner=t, lanes_per_chain=self.lanes_per_chain, thin=self.thin)
