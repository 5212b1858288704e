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

    def mc_se(self, mc_cov_mat=None, method="inse", adjust=False):
        """chain_lists.py:64-71: per-chain Monte Carlo standard errors, [C, P]."""
        cov_ = self.mc_cov(method=method, adjust=adjust) if mc_cov_mat is None else torch.stack(list(mc_cov_mat))
        return torch.diagonal(cov_, dim1=1, dim2=2).sqrt()

    def mc_se_summary(self, g=lambda x: torch.mean(x, dim=0), mc_cov_mat=None, method="inse", adjust=False):
        return g(self.mc_se(mc_cov_mat=mc_cov_mat, method=method, adjust=adjust))

    def mc_cov_summary(self, g=lambda m: torch.mean(m, dim=0), method="inse", adjust=False):
        return g(self.mc_cov(method=method, adjust=adjust))

    def multi_rhat(self, mc_cov_mat=None, method="inse", adjust=False):
        """chain_lists.py:122-123."""
        from .. import stats as st
        return st.multi_rhat(self.get_samples(), mc_cov_mat=mc_cov_mat, method=method, adjust=adjust)

    def summary(self, keys=("multi_ess", "multi_rhat"), g_mean_summary=lambda x: torch.mean(x, dim=0),
                g_mc_se_summary=lambda x: torch.mean(x, dim=0), g_acceptance_summary=lambda x: sum(x) / len(x),
                g_multi_ess_summary=lambda x: sum(x) / len(x), mc_cov_mat=None, method="inse", adjust=False):
        """chain_lists.py:125-155."""
        out = {}
        if mc_cov_mat is None and any(k in keys for k in ("mc_se", "multi_rhat")):
            mc_cov_mat = self.mc_cov(method=method, adjust=adjust)
        for key in keys:
            if key == "mean":
                out[key] = self.mean_summary(g=g_mean_summary)
            elif key == "mc_se":
                out[key] = self.mc_se_summary(g=g_mc_se_summary, mc_cov_mat=mc_cov_mat)
            elif key == "acceptance":
                out[key] = self.acceptance_summary(g=g_acceptance_summary)
            elif key == "multi_ess":
                out[key] = self.multi_ess_summary(g=g_multi_ess_summary, method=method, adjust=adjust)
            elif key == "multi_rhat":
                out[key] = self.multi_rhat(mc_cov_mat=mc_cov_mat, method=method, adjust=adjust)[0]
        return out
