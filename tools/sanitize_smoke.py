"""Small invocation of every kernel family, meant to run under `compute-sanitizer --tool memcheck` on a GPU box:
    compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py
Sizes are tiny (ragged on purpose) so that the instrumented run finishes quickly."""
import os
import sys

import numpy as np
import torch
from torch.distributions import Normal
from torch.utils.data import DataLoader

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eeyore_b200 import stats as st  # noqa: E402
from eeyore_b200.constants import loss_functions  # noqa: E402
from eeyore_b200.datasets import XYDataset  # noqa: E402
from eeyore_b200.models.mlp import MLP, Hyperparameters  # noqa: E402
from eeyore_b200.samplers import HMC, MALA, SMMALA, DataShardedHMC, MetropolisHastings  # noqa: E402


def model_for(dims, loss, dt):
    nl = len(dims) - 1
    last = torch.sigmoid if loss == "binary_classification" else None
    m = MLP(loss=loss_functions[loss], hparams=Hyperparameters(dims, nl * [True], (nl - 1) * [torch.sigmoid] + [last]), dtype=dt)
    p = m.num_params()
    m.prior = Normal(torch.zeros(p, dtype=dt), 1.7 * torch.ones(p, dtype=dt))
    return m


def main():
    g = torch.Generator().manual_seed(0)
    for dt in (torch.float64, torch.float32):
        xor = XYDataset.from_eeyore("xor", dtype=dt)
        iris = XYDataset.from_eeyore("iris", yndmin=1, yonehot=True, dtype=dt)
        for dims, loss, ds in (([2, 2, 1], "binary_classification", xor), ([2, 3, 2, 1], "binary_classification", xor),
                               ([4, 3, 3], "multiclass_classification", iris), ([4, 3, 2, 3], "multiclass_classification", iris)):
            m = model_for(dims, loss, dt)
            p = m.num_params()
            th = torch.randn(37, p, generator=g, dtype=dt) * 0.5
            for lanes in (1, 4, 8, 16, 32):
                m.upto_grad_log_target_batch(th, ds.x, ds.y, lanes=lanes)
                m.log_target_batch(th, ds.x, ds.y, lanes=lanes)
            m.forward_batch(th, ds.x)
            loader = DataLoader(ds, batch_size=len(ds))
            for cls, kw in ((MetropolisHastings, {}), (MALA, dict(step=0.01)), (HMC, dict(step=0.02, num_steps=3))):
                for lanes in (1, 8, 32):
                    s = cls(m, theta0=th, dataloader=loader, lanes_per_chain=lanes, seed=1, **kw)
                    s.run(num_epochs=4, num_burnin_epochs=1)
                    s.get_chain().get_samples().sum().item()
            if loss == "binary_classification":
                s = SMMALA(m, theta0=th, dataloader=loader, step=0.3, seed=2)
                s.run(num_epochs=3, num_burnin_epochs=0)
                s.get_chain().get_samples().sum().item()
    x = torch.randn(5, 203, 7, generator=g, dtype=torch.float64).cumsum(1) * 0.1 + torch.randn(5, 203, 7, generator=g, dtype=torch.float64)
    st.chain_stats(x, want=("mean", "cov", "inse", "ess"), max_lag=9, check=False)
    st.chain_stats(x.permute(1, 2, 0).contiguous().cuda(), layout="npc", want=("ess",), check=False)
    wide = model_for([16, 64, 64, 1], "binary_classification", torch.float32)
    xs = torch.randn(1003, 16, generator=g)
    ys = (torch.rand(1003, 1, generator=g) < 0.5).float()
    wide.upto_grad_log_target(torch.randn(5313, generator=g) * 0.1, xs, ys)
    d = DataShardedHMC(wide, torch.randn(5313, generator=g) * 0.1, xs[:1000], ys[:1000], step=1e-3, num_steps=2, seed=3)
    d.run(num_epochs=2, num_burnin_epochs=0)
    torch.cuda.synchronize()
    print("sanitize_smoke: all kernels ran")


if __name__ == "__main__":
    main()
