"""Several chains side by side; mirror of eeyore/chains/chain_lists.py:7-155 (vals[key][chain][iteration])."""
import torch

from .chain_file import ChainFile

_DEFAULT_KEYS = ("sample", "target_val", "accepted")


class ChainLists:
    def __init__(self, keys=_DEFAULT_KEYS, vals=None):
        self.reset(keys=keys, vals=vals)

    def reset(self, keys=_DEFAULT_KEYS, vals=None):
        self.vals = {key: [] for key in keys} if vals is None else vals

    @classmethod
    def from_chain_list(cls, chain_lists, keys=_DEFAULT_KEYS):
        common = set.intersection(*[set(c.vals.keys()) for c in chain_lists]) & set(keys)
        return cls(keys=common, vals={k: [c.vals[k] for c in chain_lists] for k in common})

    @classmethod
    def from_file(cls, paths, keys=_DEFAULT_KEYS, mode="a", dtype=torch.float64, device="cpu"):
        return cls.from_chain_list(
            [ChainFile(keys=keys, path=p, mode=mode).to_chainlist(dtype=dtype, device=device) for p in paths], keys=keys)

    def __repr__(self):
        return f"{len(self)} Markov chains, each containing {self.num_samples()} samples."

    def __len__(self):
        return self.num_chains()

    def num_params(self):
        return len(self.vals["sample"][0][0])

    def num_samples(self):
        return len(self.vals["sample"][0])

    def num_chains(self):
        return len(self.vals["sample"])

    def get_chain(self, idx, key="sample"):
        return torch.stack(list(self.vals[key][idx]))

    def get_samples(self):
        return torch.stack([self.get_chain(i) for i in range(self.num_chains())])

    def get_target_vals(self):
        return torch.stack([self.get_chain(i, key="target_val") for i in range(self.num_chains())])

    def get_grad_vals(self):
        return torch.stack([self.get_chain(i, key="grad_val") for i in range(self.num_chains())])

    def mean(self):
        return self.get_samples().mean(1)

    def mean_summary(self, g=lambda x: torch.mean(x, dim=0)):
        return g(self.mean())

    def acceptance(self):
        return [sum(a) / self.num_samples() for a in self.vals["accepted"]]

    def acceptance_summary(self, g=lambda x: sum(x) / len(x)):
        return g(self.acceptance())

    def mc_cov(self, method="inse", adjust=False):
        from .. import stats as st
        return st.mc_cov_batch(self.get_samples(), method=method, adjust=adjust)

    def multi_ess(self, mc_cov_mat=None, method="inse", adjust=False):
        """Per-chain multivariate ESS; all chains are processed by one device launch."""
        from .. import stats as st
        return st.multi_ess_batch(self.get_samples(), method=method, adjust=adjust).tolist()

    def multi_ess_summary(self, g=lambda x: sum(x) / len(x), mc_cov_mat=None, method="inse", adjust=False):
        return g(self.multi_ess(mc_cov_mat=mc_cov_mat, method=method, adjust=adjust))
