"""CSV -> tensors; mirror of eeyore/datasets/xydataset.py:11-53 (host side, kept in Python)."""
from pathlib import Path

import numpy as np
import torch
from torch.nn.functional import one_hot
from torch.utils.data import Dataset

from ..constants import torch_to_np_types
from .data_info import data_paths


def _read(path, dtype, skiprows, usecols, ndmin, onehot, device):
    t = torch.from_numpy(np.loadtxt(path, dtype=torch_to_np_types[dtype], delimiter=",", skiprows=skiprows,
                                    usecols=usecols, ndmin=ndmin)).to(device=device)
    return one_hot(t.long()).to(t.dtype) if onehot else t


class XYDataset(Dataset):
    def __init__(self, x, y):
        self.set_data(x, y)

    def __repr__(self):
        return "XYDataset"

    def __len__(self):
        return len(self.x)

    def __getitem__(self, idx):
        return self.x[idx], self.y[idx]

    def set_data(self, x, y):
        self.x, self.y = x, y

    @classmethod
    def from_file(cls, path=Path.cwd(), xfile="x.csv", yfile="y.csv", xskiprows=1, yskiprows=1, xusecols=None,
                  yusecols=None, xndmin=2, yndmin=2, dtype=torch.float64, device="cpu", xonehot=False,
                  yonehot=False):
        path = Path(path)
        return cls(_read(path / xfile, dtype, xskiprows, xusecols, xndmin, xonehot, device),
                   _read(path / yfile, dtype, yskiprows, yusecols, yndmin, yonehot, device))

    @classmethod
    def from_eeyore(cls, data_name, xndmin=2, yndmin=2, dtype=torch.float64, device="cpu", xonehot=False,
                    yonehot=False):
        return cls.from_file(path=data_paths[data_name], xndmin=xndmin, yndmin=yndmin, dtype=dtype, device=device,
                             xonehot=xonehot, yonehot=yonehot)
