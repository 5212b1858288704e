"""Times the UNMODIFIED reference (/root/reference, read-only) on this container's CPU for the BASELINE shapes.
Build-container only (the reference cannot travel to the GPU box); the output is committed under profiles/ and quoted by
bench.py in cpu_baseline.note.   python tools/time_reference.py > profiles/r02_reference_cpu_timing.json"""
import json
import math
import os
import platform
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle" / "kanga_stub"))
sys.path.insert(0, "/root/reference")

from torch.distributions import Normal  # noqa: E402
from torch.utils.data import DataLoader  # noqa: E402

from eeyore.constants import loss_functions  # noqa: E402
from eeyore.datasets import XYDataset  # noqa: E402
from eeyore.models.mlp import MLP, Hyperparameters  # noqa: E402
from eeyore.samplers import HMC, MALA  # noqa: E402
import eeyore.stats as st  # noqa: E402

torch.set_num_threads(1)
S3 = math.sqrt(3.0)


def model_of(dims, loss, dtype=torch.float64):
    nl = len(dims) - 1
    last = torch.sigmoid if loss == "binary_classification" else None
    m = MLP(loss=loss_functions[loss], hparams=Hyperparameters(dims, nl * [True], (nl - 1) * [torch.sigmoid] + [last]), dtype=dtype)
    p = m.num_params()
    m.prior = Normal(torch.zeros(p, dtype=dtype), S3 * torch.ones(p, dtype=dtype))
    return m


def timed(fn):
    t0 = time.perf_counter()
    fn()
    return time.perf_counter() - t0


out = {"host": {"cpu": platform.processor() or platform.machine(), "cores": os.cpu_count(), "torch": torch.__version__,
                "torch_threads": 1}}
xor = XYDataset.from_eeyore("xor", dtype=torch.float64)
# config 1: exactly MALA.run(1100, 110), MLP 2-2-1 on XOR
torch.manual_seed(0)
m = model_of([2, 2, 1], "binary_classification")
s = MALA(m, theta0=m.prior.sample(), dataloader=DataLoader(xor, batch_size=4), step=1.74)
t = timed(lambda: s.run(num_epochs=1100, num_burnin_epochs=110))
out["cfg1_mala_221_1100_iters"] = {"seconds": t, "evals_per_s": 1100 / t}
# config 4 shape: HMC L = 10, MLP 2-3-2-1 on XOR, one chain, 100 iterations (the reference executes L + 1 evaluations)
m = model_of([2, 3, 2, 1], "binary_classification")
s = HMC(m, theta0=m.prior.sample(), dataloader=DataLoader(xor, batch_size=4), step=0.3, num_steps=10)
t = timed(lambda: s.run(num_epochs=100, num_burnin_epochs=0))
out["cfg4_hmc_2321_xor"] = {"seconds": t, "iterations": 100, "evals_per_s_counting_L": 1000 / t,
                            "evals_per_s_as_executed_L_plus_1": 1100 / t}
# config 2 shape: HMC L = 10, MLP 4-3-3, N = 150
rng = np.random.default_rng(1)
centres = rng.normal(size=(3, 4)) * 2.0
x = np.concatenate([centres[k] + 0.5 * rng.normal(size=(50, 4)) for k in range(3)])
y = np.eye(3)[np.repeat(np.arange(3), 50)]
ds = XYDataset(torch.from_numpy(x), torch.from_numpy(y))
m = model_of([4, 3, 3], "multiclass_classification")
s = HMC(m, theta0=m.prior.sample() * 0.3, dataloader=DataLoader(ds, batch_size=150), step=0.15, num_steps=10)
t = timed(lambda: s.run(num_epochs=60, num_burnin_epochs=0))
out["cfg2_hmc_433_n150"] = {"seconds": t, "iterations": 60, "evals_per_s_counting_L": 600 / t}
# config 3 shape: one log-target + gradient evaluation, MLP 2-3-2-1, N = 200 (SMMALA itself is absent from the reference)
rng = np.random.default_rng(3)
corners = np.array([[0, 0], [0, 1], [1, 0], [1, 1]], dtype=np.float64)
x = np.concatenate([c + 0.15 * rng.normal(size=(50, 2)) for c in corners])
y = np.concatenate([np.full((50, 1), float(int(c[0]) ^ int(c[1]))) for c in corners])
m = model_of([2, 3, 2, 1], "binary_classification")
th = m.prior.sample() * 0.5
xt, yt = torch.from_numpy(x), torch.from_numpy(y)
t = timed(lambda: [m.upto_grad_log_target(th.clone().detach(), xt, yt) for _ in range(300)])
out["cfg3_upto_grad_2321_n200"] = {"seconds_per_eval": t / 300, "evals_per_s": 300 / t}
# config 5 shape: 16-64-64-1, fp32, N = 1,048,576 rows (1/8 of the data set), all threads
torch.set_num_threads(os.cpu_count())
m = model_of([16, 64, 64, 1], "binary_classification", torch.float32)
m.prior = Normal(torch.zeros(m.num_params()), S3 * torch.ones(m.num_params()))
g = torch.Generator().manual_seed(4)
xt = torch.randn(1 << 20, 16, generator=g)
yt = (torch.rand(1 << 20, 1, generator=g) < 0.5).float()
th = torch.randn(m.num_params(), generator=g) * 0.1
m.upto_grad_log_target(th.clone().detach(), xt, yt)
t = timed(lambda: [m.upto_grad_log_target(th.clone().detach(), xt, yt) for _ in range(3)]) / 3
out["cfg5_upto_grad_16_64_64_1_n1M"] = {"seconds_per_eval_1M_rows": t, "torch_threads": os.cpu_count(),
                                         "evals_per_s_scaled_to_8388608_rows": 1.0 / (8 * t)}
torch.set_num_threads(1)
# diagnostics: multi_ess of one [1000, 20] chain
ch = torch.randn(1000, 20, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
for i in range(1, 1000):
    ch[i] = 0.3 * ch[i - 1] + ch[i]
t = timed(lambda: st.multi_ess(ch))
out["multi_ess_1000x20"] = {"seconds_per_chain": t}
print(json.dumps(out, indent=1))
