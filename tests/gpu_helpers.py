"""Helpers for the -m gpu parity tests (call the product through its Python mirror / C ABI)."""
import numpy as np
import torch
from torch.distributions import Normal
from torch.utils.data import DataLoader

from eeyore_b200.constants import loss_functions
from eeyore_b200.datasets import XYDataset
from eeyore_b200.models.mlp import MLP, Hyperparameters
from helpers import ARCHS, load, data_of

T_DTYPES = {"f64": torch.float64, "f32": torch.float32}


def make_model(arch, tag, prior_scale=1.0, temperature=None):
    a = ARCHS[arch]
    dims = a["dims"]
    nl = len(dims) - 1
    binary = a["loss"] == "binary_classification"
    hp = Hyperparameters(dims, nl * [True], (nl - 1) * [torch.sigmoid] + [torch.sigmoid if binary else None])
    dt = T_DTYPES[tag]
    m = MLP(loss=loss_functions[a["loss"]], hparams=hp, dtype=dt, temperature=temperature)
    p = m.num_params()
    m.prior = Normal(torch.zeros(p, dtype=dt), prior_scale * torch.ones(p, dtype=dt))
    return m


def dataset(arch, tag):
    dt = T_DTYPES[tag]
    if ARCHS[arch]["data"] == "xor":
        return XYDataset.from_eeyore("xor", dtype=dt)
    return XYDataset.from_eeyore("iris", yndmin=1, yonehot=True, dtype=dt)


def loader(ds):
    return DataLoader(ds, batch_size=len(ds))


def npy(t):
    return t.detach().cpu().numpy()
