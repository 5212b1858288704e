"""Mirror of eeyore/linalg/is_pos_def.py:3-11.  Inside the device diagnostics the same test ("exactly symmetric and
Cholesky succeeds") is the warp-level factorisation of eeyore_b200/csrc/stats.cu; this host helper keeps the API."""
import torch


def is_pos_def(x):
    if not torch.equal(x, x.t()):
        return False
    return bool(torch.linalg.cholesky_ex(x).info.item() == 0)
