"""Saved states of C chains kept on the device (no reference counterpart; it is the chain-batched analogue of the
per-run ChainList objects SerialSampler.benchmark produces, eeyore/samplers/serial_sampler.py:76-94).

Storage is structure-of-arrays so that both the sampler's stores and the diagnostics' loads are coalesced:
  samples   [n, P, C]   (element (s, c, j) at s*P*C + j*C + c)
  grad_vals [n, P, C]
  target    [n, C]
  accepted  [n, C] uint8
``get_samples()`` returns the reference's [C, n, P] orientation as a strided view.
"""
from pathlib import Path

import torch


class DeviceChains:
    def __init__(self, samples, target_vals=None, grad_vals=None, accepted=None, accept_count=None, n_iters=None):
        self.samples_soa = samples
        self.target_soa = target_vals
        self.grad_soa = grad_vals
        self.accepted_soa = accepted
        self.accept_count = accept_count
        self.n_iters = n_iters

    def __repr__(self):
        return f"{self.num_chains()} Markov chains on {self.samples_soa.device}, each containing {self.num_samples()} samples."

    def __len__(self):
        return self.num_chains()

    def num_chains(self):
        return self.samples_soa.shape[2]

    def num_samples(self):
        return self.samples_soa.shape[0]

    def num_params(self):
        return self.samples_soa.shape[1]

    def get_samples(self):
        """[C, n, P] view (ChainLists.get_samples orientation, chain_lists.py:52-53)."""
        return self.samples_soa.permute(2, 0, 1)

    def get_chain(self, idx, key="sample"):
        if key == "sample":
            return self.samples_soa[:, :, idx]
        if key == "grad_val":
            return self.grad_soa[:, :, idx]
        if key == "target_val":
            return self.target_soa[:, idx]
        if key == "accepted":
            return self.accepted_soa[:, idx]
        raise KeyError(key)

    def get_target_vals(self):
        return self.target_soa.t()

    def get_grad_vals(self):
        return self.grad_soa.permute(2, 0, 1)

    def mean(self):
        return self.samples_soa.mean(0).t()

    def acceptance(self):
        """Per-chain acceptance rate over the saved iterations."""
        return self.accepted_soa.to(torch.float64).mean(0)

    def acceptance_summary(self, g=lambda x: x.mean().item()):
        return g(self.acceptance())

    def multi_ess(self):
        from .. import stats as st
        return st.multi_ess_soa(self.samples_soa)

    def multi_rhat(self):
        """Across-chain multivariate R-hat (eeyore/stats/multi_rhat.py) of the saved samples."""
        from .. import stats as st
        return st.multi_rhat(self.get_samples())

    def acf(self, max_lag):
        from .. import stats as st
        return st.acf_soa(self.samples_soa, max_lag)

    def to_chainlist(self, idx):
        from .chain_list import ChainList
        vals = {"sample": list(self.get_chain(idx).unbind(0))}
        if self.target_soa is not None:
            vals["target_val"] = list(self.target_soa[:, idx].unbind(0))
        if self.grad_soa is not None:
            vals["grad_val"] = list(self.grad_soa[:, :, idx].unbind(0))
        if self.accepted_soa is not None:
            vals["accepted"] = [int(a) for a in self.accepted_soa[:, idx].tolist()]
        return ChainList(vals=vals)

    def to_chainlists(self, keys=("sample", "target_val", "accepted")):
        from .chain_lists import ChainLists
        return ChainLists.from_chain_list([self.to_chainlist(i) for i in range(self.num_chains())], keys=keys)

    def to_chainfiles(self, path, keys=("sample", "target_val", "accepted"), mode="w"):
        """Writes the run%0Nd/<key>.csv tree of SerialSampler.benchmark (serial_sampler.py:76-94)."""
        from .chain_file import ChainFile
        c = self.num_chains()
        host = {"sample": self.get_samples().cpu()}
        if self.target_soa is not None:
            host["target_val"] = self.get_target_vals().cpu()
        if self.grad_soa is not None:
            host["grad_val"] = self.get_grad_vals().cpu()
        if self.accepted_soa is not None:
            host["accepted"] = self.accepted_soa.t().cpu()
        for i in range(c):
            run = Path(path) / ("run" + str(i + 1).zfill(len(str(c))))
            cf = ChainFile(keys=[k for k in keys if k in host], path=run, mode=mode)
            cf.write_block({k: (host[k][i].tolist() if k == "accepted" else host[k][i]) for k in cf.vals.keys()})
            cf.close()
