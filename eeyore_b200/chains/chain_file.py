"""Chain stored as one CSV file per key; mirror of eeyore/chains/chain_file.py:9-81 (output boundary, host side).

File format (chain_file.py:28-45): `<key>.csv`, one row per saved iteration, '%.18e' comma-separated for tensors,
a bare integer per line for 'accepted'.
"""
from pathlib import Path

import numpy as np
import torch

from ..constants import torch_to_np_types
from .chain import Chain

_DEFAULT_KEYS = ("sample", "target_val", "accepted")
_DEFAULT_FMT = {"sample": "%.18e", "target_val": "%.18e", "grad_val": "%.18e", "accepted": "%d"}


class ChainFile(Chain):
    def __init__(self, keys=_DEFAULT_KEYS, path=Path.cwd(), mode="a"):
        self.path = Path(path)
        self.mode = mode
        self.path.mkdir(parents=True, exist_ok=True)
        self.reset(keys=keys)

    def reset(self, keys=_DEFAULT_KEYS):
        self.vals = {key: open(self.path / f"{key}.csv", self.mode) for key in keys}

    def close(self):
        for f in self.vals.values():
            f.close()

    def update(self, state, reset=True, close=True, fmt=_DEFAULT_FMT):
        if reset:
            self.reset(keys=list(self.vals.keys()))
        self.write_block({k: [state[k]] for k in self.vals.keys()}, fmt=fmt)
        if close:
            self.close()

    def write_block(self, block, fmt=_DEFAULT_FMT):
        """block[key] = sequence (or [n, ...] tensor) of per-iteration values; written with one savetxt per key."""
        for key, f in self.vals.items():
            vals = block[key]
            if isinstance(vals, torch.Tensor):
                arr = vals.detach().cpu().numpy()
            elif len(vals) and isinstance(vals[0], torch.Tensor):
                arr = torch.stack(list(vals)).detach().cpu().numpy()
            elif len(vals) and isinstance(vals[0], np.ndarray):
                arr = np.stack(vals)
            else:
                f.write("".join(f"{v}\n" for v in vals))
                continue
            if arr.shape[0]:
                np.savetxt(f, arr.reshape(arr.shape[0], -1), fmt=fmt[key], delimiter=",")

    def extend_from_device(self, samples=None, target_vals=None, grad_vals=None, accepted=None, fmt=_DEFAULT_FMT):
        """Append the saved states of one fused sampler launch ([n, P], [n], [n, P], [n] device tensors) to the files --
        what n consecutive update() calls of the reference's loop (chain_file.py:28-45) would have written."""
        block = {"sample": samples, "target_val": target_vals, "grad_val": grad_vals,
                 "accepted": None if accepted is None else [int(a) for a in accepted.tolist()]}
        missing = [k for k in self.vals if block.get(k) is None]
        if missing:
            raise KeyError(f"the native samplers do not record {missing}")
        if any(f.closed for f in self.vals.values()):
            self.reset(keys=list(self.vals.keys()))
        self.write_block(block, fmt=fmt)
        self.close()

    def line_to_val_element(self, line, key, dtype=torch.float64, device="cpu"):
        if key == "accepted":
            return int(line.strip())
        np_t = torch_to_np_types[dtype]
        if key == "target_val":
            return torch.tensor(np_t(line.strip())).to(device=device)
        return torch.tensor([np_t(v) for v in line.split(",")]).to(device=device)

    def to_chainlist(self, keys=None, dtype=torch.float64, device="cpu"):
        from .chain_list import ChainList
        keys = [k for k in (keys or self.vals.keys()) if k in ("sample", "target_val", "grad_val", "accepted")]
        vals = {}
        for key in keys:
            with open(self.path / f"{key}.csv") as f:
                vals[key] = [self.line_to_val_element(line, key, dtype=dtype, device=device) for line in f]
        return ChainList(vals=vals)
