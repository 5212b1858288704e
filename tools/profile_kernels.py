"""Runs the SMMALA (cfg3-shaped) and chain-statistics kernels once each at benchmark-like sizes (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.distributions import Normal
from torch.utils.data import DataLoader
from eeyore_b200 import stats as st
from eeyore_b200.constants import loss_functions
from eeyore_b200.datasets import XYDataset
from eeyore_b200.models.mlp import MLP, Hyperparameters
from eeyore_b200.samplers import HMC, SMMALA

rng = np.random.default_rng(3)
corners = np.array([[0, 0], [0, 1], [1, 0], [1, 1]], dtype=np.float64)
x = np.concatenate([c + 0.15 * rng.normal(size=(50, 2)) for c in corners])
y = np.concatenate([np.full((50, 1), float(int(c[0]) ^ int(c[1]))) for c in corners])
ds = XYDataset(torch.from_numpy(x), torch.from_numpy(y))
m = MLP(loss=loss_functions["binary_classification"], hparams=Hyperparameters([2, 3, 2, 1], 3 * [True], 3 * [torch.sigmoid]))
m.prior = Normal(torch.zeros(20, dtype=torch.float64), 3 ** 0.5 * torch.ones(20, dtype=torch.float64))
C = 16384
th = torch.randn(C, 20, dtype=torch.float64, generator=torch.Generator().manual_seed(0)) * 0.5
for _ in range(2):
    s = SMMALA(m, theta0=th, dataloader=DataLoader(ds, batch_size=200), step=0.3, seed=1)
    s.run(num_epochs=10, num_burnin_epochs=0)
torch.cuda.synchronize()
print("smmala acceptance", s.get_chain().acceptance().mean().item())
# chain statistics: 4096 chains x 1000 samples x 20 parameters from an HMC run on XOR
xor = XYDataset.from_eeyore("xor")
h = HMC(m, theta0=torch.randn(4096, 20, dtype=torch.float64, generator=torch.Generator().manual_seed(1)),
        dataloader=DataLoader(xor, batch_size=4), step=0.3, num_steps=10, seed=2)
h.run(num_epochs=1100, num_burnin_epochs=100)
soa = h.get_chain().samples_soa
for _ in range(2):
    out = st.chain_stats(soa, layout="npc", want=("ess",), max_lag=50, check=False)
torch.cuda.synchronize()
print("stats ok", float((out["status"] == 0).double().mean()), float(out["ess"][out["status"] == 0].mean()))
