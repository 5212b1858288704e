"""In-memory Monte Carlo chain; mirror of eeyore/chains/chain_list.py:12-141 (output boundary, host side).

``extend_from_device`` is the native fast path: a whole run's saved states arrive as device buffers and are
appended in one go instead of one python ``update`` per iteration.  Diagnostics (mc_cov / multi_ess) run on the
device through eeyore_b200.stats.
"""
from pathlib import Path

import numpy as np
import torch

from .chain import Chain

_DEFAULT_KEYS = ("sample", "target_val", "accepted")
_DEFAULT_FMT = {"sample": "%.18e", "target_val": "%.18e", "grad_val": "%.18e", "accepted": "%d"}


class ChainList(Chain):
    def __init__(self, keys=_DEFAULT_KEYS, vals=None):
        self.reset(keys=keys, vals=vals)

    def reset(self, keys=_DEFAULT_KEYS, vals=None):
        self.vals = {key: [] for key in keys} if vals is None else vals

    def __repr__(self):
        return f"Markov chain containing {len(self)} samples."

    def __len__(self):
        return self.num_samples()

    def num_params(self):
        return len(self.get_sample(0))

    def num_samples(self):
        return len(self.vals["sample"])

    def get_param(self, idx):
        return torch.stack([s[idx] for s in self.vals["sample"]])

    def get_sample(self, idx):
        return self.vals["sample"][idx]

    def get_samples(self):
        return torch.stack(self.vals["sample"])

    def get_target_vals(self):
        return torch.stack(self.vals["target_val"])

    def get_grad_val(self, idx):
        return self.vals["grad_val"][idx]

    def get_grad_vals(self):
        return torch.stack(self.vals["grad_val"])

    def state(self, idx=-1):
        current = {}
        for key, val in self.vals.items():
            try:
                current[key] = val[idx]
            except IndexError:
                print(f"WARNING: chain does not have values for {key}.")
        return current

    def update(self, state):
        for key in self.vals.keys():
            self.vals[key].append(state[key])

    def extend_from_device(self, samples=None, target_vals=None, grad_vals=None, accepted=None):
        """Append a block of saved states ([n, P], [n], [n, P], [n]) produced by one fused sampler launch."""
        block = {"sample": samples, "target_val": target_vals, "grad_val": grad_vals}
        for key in self.vals.keys():
            if key == "accepted":
                self.vals[key].extend(int(a) for a in accepted.tolist())
            elif key in block and block[key] is not None:
                self.vals[key].extend(block[key].unbind(0))
            else:
                raise KeyError(f"the native samplers do not record '{key}'")

    def mean(self):
        return self.get_samples().mean(0)

    def running_mean(self, idx):
        from .. import stats as st
        return st.running_mean(self.get_param(idx))

    def running_means(self):
        from .. import stats as st
        return st.running_mean(self.get_samples(), dim=0)

    def mc_cov(self, method="inse", adjust=False):
        from .. import stats as st
        return st.mc_cov(self.get_samples(), method=method, adjust=adjust, rowvar=False)

    def mc_se(self, mc_cov_mat=None, method="inse", adjust=False):
        from .. import stats as st
        if mc_cov_mat is None:
            return st.mc_se(self.get_samples(), method=method, adjust=adjust, rowvar=False)
        return st.mc_se_from_cov(mc_cov_mat)

    def acceptance_rate(self):
        return sum(self.vals["accepted"]) / self.num_samples()

    def multi_ess(self, mc_cov_mat=None, method="inse", adjust=False):
        from .. import stats as st
        return st.multi_ess(self.get_samples(), mc_cov_mat=mc_cov_mat, method=method, adjust=adjust)

    def save(self, path):
        torch.save(self.vals, path)

    def load(self, path):
        self.vals = torch.load(path)

    def to_chainfile(self, keys=None, path=Path.cwd(), mode="a", fmt=_DEFAULT_FMT):
        from .chain_file import ChainFile
        cf = ChainFile(keys=keys or self.vals.keys(), path=Path(path), mode=mode)
        cf.write_block({k: self.vals[k] for k in cf.vals.keys()}, fmt=fmt)
        cf.close()
