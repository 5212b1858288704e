"""Power-posterior sampler: K tempered chains (MH or MALA within a level) with neighbour swaps between levels.
Mirror of eeyore/samplers/power_posterior_sampler.py:15-182 (and multi_chain_serial_sampler.py:5-46), batched over E
independent ensembles: theta0 [P] is the reference's single ensemble, theta0 [E, P] runs E of them side by side.

Device path, no host synchronisation inside run():
  * within-chain moves: one fused launch per level for all iterations up to the next between-chain sweep
    (eeyore_b200_mh_run / eeyore_b200_mala_run over the level's E chains at the level's temperature);
  * between-chain sweep (power_posterior_sampler.py:136-171), sequential in the level i as in the reference and vectorised
    over the ensembles: two batched log-likelihood / log-prior / gradient evaluations (eeyore_b200_log_target_grad at
    theta_j and at theta_i, temperature applied per ensemble as T*loglik + T*logprior, bayesian_model.py:30-56), the
    categorical proposal terms, the accept test and the state exchange as element-wise device operations.
The state saved for an iteration that ends with a sweep is the post-sweep state (power_posterior_sampler.py:173-182).
Reference behaviour kept on purpose: the samplers' initial target / gradient are evaluated before the temperature ladder
is assigned (power_posterior_sampler.py:33-35), i.e. untempered, and stay so until the level's first accepted move.
Full-batch iterations only (one batch per epoch); thinning is not offered (the reference has none).
"""
import copy
import math
from pathlib import Path

import torch

from .. import _native as nv
from ..chains import ChainFile, ChainList
from ..datasets import DataCounter
from .mala import MALA
from .metropolis_hastings import MetropolisHastings
from .serial_sampler import SerialSampler


class PowerPosteriorSampler(SerialSampler):
    def __init__(self, model, dataloader, samplers, theta0=None, data0=None, counter=None, temperature=None,
                 between_step=10, b=0.5, storage="list", keys=("sample", "target_val"), path=Path.cwd(), mode="a",
                 check_input=False, seed=None):
        super().__init__(counter or DataCounter.from_dataloader(dataloader))
        if theta0 is None:
            raise ValueError("the device power-posterior sampler needs theta0 ([P] or [E, P])")
        self.between_step, self.b = int(between_step), float(b)
        self.num_chains = len(samplers)
        if self.num_chains < 2:
            raise ValueError("a power-posterior ladder needs at least two levels")
        self.dataloader = dataloader
        self.sampler_names = [samplers[i][0] for i in range(self.num_chains)]
        self.keys = list(keys)
        self.seed = int(seed) if seed is not None else int(torch.randint(0, 2 ** 62, (1,)).item())
        self._tape = None
        self.init_samplers(model, samplers, theta0, data0 or next(iter(dataloader)), storage, self.keys, path, mode)
        self.set_temperature(temperature)          # after the initial evaluation, as in the reference (:33-35)
        self.dtype, self.device = self.samplers[0].model.dtype, self.samplers[0]._theta.device
        if check_input:
            self.check_dtype()
        self._raw = copy.deepcopy(model)           # untempered model for the cross evaluations of the sweeps
        self._raw.temperature = None
        self._needs_grad = any(s._uses_grad for s in self.samplers)
        lq = self._categorical_log_probs()
        self._lq = torch.tensor(lq, dtype=torch.float64, device=self.device)
        cdf = torch.tensor([[math.exp(v) if v > -math.inf else 0.0 for v in row] for row in lq], dtype=torch.float64)
        self._cdf = torch.cumsum(cdf, dim=1).to(self.device)
        self._gen = torch.Generator(device=self.device).manual_seed(self.seed)
        self.num_between_sweeps = 0
        self.swap_count = torch.zeros(self.num_chains, dtype=torch.int64, device=self.device)

    # -- construction (power_posterior_sampler.py:45-97) ----------------------------------------------------------------
    def check_dtype(self):
        if not all(s.model.dtype == self.dtype for s in self.samplers):
            raise ValueError

    def init_chain(self, i, storage, keys, path, mode):
        if storage == "list":
            return ChainList(keys=list(keys))
        if storage == "file":
            chain_path = Path(path).joinpath("chain" + f"{(i + 1):0{len(str(self.num_chains))}}")
            chain_path.mkdir(parents=True, exist_ok=True)
            return ChainFile(keys=list(keys), path=chain_path, mode=mode)
        raise ValueError("storage must be 'list' or 'file'")

    def init_samplers(self, model, samplers, theta0, data0, storage, keys, path, mode):
        self.samplers = []
        for i, (name, kwargs) in enumerate(samplers):
            cls = {"MetropolisHastings": MetropolisHastings, "MALA": MALA}.get(name)
            if cls is None:
                raise ValueError("within-chain samplers are 'MetropolisHastings' or 'MALA' (power_posterior_sampler.py:69-84)")
            m = copy.deepcopy(model)
            m.temperature = None
            self.samplers.append(cls(m, theta0=theta0, dataloader=self.dataloader, data0=data0,
                                     chain=self.init_chain(i, storage, keys, path, mode), seed=self.seed + 7919 * (i + 1),
                                     **kwargs))
        self.num_ensembles = self.samplers[0].num_chains
        self._batched = self.samplers[0]._batched

    def default_indicator(self):
        return self.num_chains - 1

    def set_temperature(self, temperature):
        if temperature is not None and self.num_chains != len(temperature):
            raise ValueError
        k = self.num_chains
        self.temperature = [(i / k) ** 4 for i in range(1, k + 1)] if temperature is None else [float(t) for t in temperature]
        for s, t in zip(self.samplers, self.temperature):
            s.model.temperature = t

    def eval_categorical_prob(self, j, i):
        eb = math.exp(-self.b)
        return eb ** abs(j - i) / (eb * (2 - eb ** i - eb ** (self.num_chains - 1 - i)) / (1 - eb))

    def _categorical_log_probs(self):
        """lq[i][j] = log-probability of proposing level j from level i: eval_categorical_probs (:113-117) through
        torch.distributions.Categorical (normalise, clamp to [eps, 1 - eps], log)."""
        k = self.num_chains
        eps = torch.finfo(torch.float64).eps
        lq = [[-math.inf] * k for _ in range(k)]
        for i in range(k):
            js = [j for j in range(k) if j != i]
            p = torch.tensor([self.eval_categorical_prob(j, i) for j in js], dtype=torch.float64)
            lp = torch.log((p / p.sum()).clamp(min=eps, max=1 - eps))
            for j, v in zip(js, lp.tolist()):
                lq[i][j] = v
        return lq

    # -- multi_chain_serial_sampler.py:10-46 ------------------------------------------------------------------------------
    def get_model(self, idx=None):
        return self.samplers[idx or self.default_indicator()].model

    def get_chain(self, idx=None):
        return self.samplers[idx or self.default_indicator()].get_chain()

    def get_param(self, param_idx, chain_idx=None):
        return self.get_chain(idx=chain_idx).get_param(param_idx)

    def get_sample(self, sample_idx, chain_idx=None):
        return self.get_chain(idx=chain_idx).get_sample(sample_idx)

    def reset(self, theta, data=None, reset_counter=True, reset_chain=True):
        x, y = data or next(iter(self.dataloader))
        for s in self.samplers:
            s.reset(theta, data=(x, y), reset_counter=reset_counter, reset_chain=reset_chain)
        self.counter.reset()

    def to_chainfile(self, path=Path.cwd(), mode="a"):
        for i, s in enumerate(self.samplers):
            s.get_chain().to_chainfile(path=Path(path).joinpath("sampler" + str(i).zfill(self.num_chains)), mode=mode)

    # -- parity mode ------------------------------------------------------------------------------------------------------
    def set_noise_tape(self, z, u, j_tape, u_between):
        """z [T, K, (E,) P] and u [T, K(, E)] for the within-chain moves, j_tape / u_between [NB, K(, E)] for the sweeps --
        the reference's draws in call order (oracle/make_golden.py: power_posterior_goldens)."""
        k, e = self.num_chains, self.num_ensembles
        z = torch.as_tensor(z).reshape(-1, k, e, self.samplers[0].model.num_params())
        u = torch.as_tensor(u).reshape(-1, k, e)
        for m, s in enumerate(self.samplers):
            s.set_noise_tape(z[:, m], u[:, m])
        self._tape = [torch.as_tensor(j_tape).reshape(-1, k, e).to(self.device, torch.int64),
                      torch.as_tensor(u_between).reshape(-1, k, e).to(self.device, self.dtype), 0]

    # -- the sweep (power_posterior_sampler.py:136-171) ---------------------------------------------------------------------
    def _draw_neighbours(self, i):
        e = self.num_ensembles
        if self._tape is not None:
            jt, ub, pos = self._tape
            if pos >= jt.shape[0]:
                raise RuntimeError("between-chain tape exhausted")
            return jt[pos, i], ub[pos, i]
        r = torch.rand(2, e, dtype=torch.float64, device=self.device, generator=self._gen)
        j = torch.searchsorted(self._cdf[i].contiguous(), r[0].clamp(max=1 - 1e-16).contiguous(), right=True)
        j = j.clamp(max=self.num_chains - 1)
        j = torch.where(j == i, torch.full_like(j, min(i + 1, self.num_chains - 1) if i < self.num_chains - 1 else i - 1), j)
        return j, r[1].to(self.dtype)

    def between_chain_moves(self, xd, yd):
        k, e = self.num_chains, self.num_ensembles
        ar = torch.arange(e, device=self.device)
        temps = torch.tensor(self.temperature, dtype=self.dtype, device=self.device)
        th = torch.stack([s._theta for s in self.samplers]).contiguous()          # [K, E, P]
        lt = torch.stack([s._lt for s in self.samplers])                          # [K, E]
        g = torch.stack([(s._grad if s._uses_grad else torch.zeros_like(s._theta)) for s in self.samplers]).contiguous() \
            if self._needs_grad else None
        for i in range(k):
            j, u = self._draw_neighbours(i)
            th_i, th_j = th[i].clone(), th[j, ar]
            lt_i, lt_j = lt[i].clone(), lt[j, ar]
            t_i, t_j = temps[i], temps[j]
            _, gr_j, ll_j, lp_j = self._raw._eval(th_j.contiguous(), xd, yd, want_grad=self._needs_grad, parts=True)
            _, gr_i, ll_i, lp_i = self._raw._eval(th_i.contiguous(), xd, yd, want_grad=self._needs_grad, parts=True)
            cross_i = t_i * ll_j + t_i * lp_j                                    # sampler_i.model.log_target(theta_j)
            cross_j = t_j * ll_i + t_j * lp_i                                    # sampler_j.model.log_target(theta_i)
            log_rate = self._lq[j, i] - self._lq[i, j] - lt_i - lt_j + cross_i + cross_j
            acc = torch.log(u) < log_rate                                        # NaN compares False -> revert
            th[i] = torch.where(acc[:, None], th_j, th_i)
            lt[i] = torch.where(acc, cross_i, lt_i)
            if g is not None:
                g_i = g[i].clone()
                g[i] = torch.where(acc[:, None], t_i * gr_j, g_i)
            for m in range(k):
                if m == i:
                    continue
                hit = acc & (j == m)
                th[m] = torch.where(hit[:, None], th_i, th[m])
                lt[m] = torch.where(hit, cross_j, lt[m])
                if g is not None:
                    g[m] = torch.where(hit[:, None], t_j[:, None] * gr_i, g[m])
            self.swap_count[i] += acc.sum()
        for m, s in enumerate(self.samplers):                                     # back into the levels' device state
            s._theta_soa.copy_(th[m].t())
            s._lt.copy_(lt[m])
            if s._uses_grad:
                s._grad_soa.copy_(g[m].t())
        if self._tape is not None:
            self._tape[2] += 1
        self.num_between_sweeps += 1

    # -- serial_sampler.py:35-52 with the draw of power_posterior_sampler.py:173-182 -----------------------------------------
    def run(self, num_epochs, num_burnin_epochs, verbose=False, verbose_step=100):
        nv.require_cuda()
        c = self.counter
        c.set_epoch_info(num_epochs, num_burnin_epochs)
        if c.num_batches != 1:
            raise ValueError("the device power-posterior sampler runs full-batch iterations (one batch per epoch)")
        if any(s.thin != 1 for s in self.samplers):
            raise ValueError("thinning is not available in the power-posterior sampler")
        xd, yd = self.samplers[0]._data_dev
        bs, end = self.between_step, c.idx + c.num_iters      # num_iters MORE draws per call, as the reference's loop
        want = tuple(self.keys)
        while c.idx < end:
            idx = c.idx
            nxt = idx if idx % bs == 0 else (idx // bs + 1) * bs     # next iteration that is followed by a sweep
            n = min(end, nxt + 1) - idx
            n_burn = max(0, min(n, c.num_burnin_iters - idx))
            outs = [s._launch(n, n_burn, xd, yd, want=want) for s in self.samplers]
            last = idx + n - 1
            if last % bs == 0:
                self.between_chain_moves(xd, yd)
                if last >= c.num_burnin_iters:                        # the saved state of `last` is the post-sweep state
                    for s, out in zip(self.samplers, outs):
                        if "sample" in out:
                            out["sample"][-1].copy_(s._theta_soa)
                        if "target_val" in out:
                            out["target_val"][-1].copy_(s._lt)
                        if "grad_val" in out:
                            out["grad_val"][-1].copy_(s._grad_soa)
            for s, out in zip(self.samplers, outs):
                s._store(out)
                if "accepted" in out:
                    s._last_accepted = out["accepted"][-1].to(torch.int64)
                s._publish_current()
            c.increment_idx(n)

    def swap_rates(self):
        """Accepted between-chain proposals per level and ensemble-sweep (diagnostic; not in the reference)."""
        return self.swap_count.to(torch.float64) / max(1, self.num_between_sweeps * self.num_ensembles)
