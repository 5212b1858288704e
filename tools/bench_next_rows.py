"""Throughput of the SURVEY 8f row-4 samplers on one GPU (log-target evaluations per second; CUDA events, 3 warm-ups):
   AM / RAM: 16,384 chains x 200 iterations of MLP 2-3-2-1 on XOR (fp64);
   PowerPosteriorSampler: 4,096 ensembles x 5 MALA levels x 200 iterations, a sweep every 10."""
import sys
from pathlib import Path

import torch
from torch.distributions import Normal
from torch.utils.data import DataLoader

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from eeyore_b200.constants import loss_functions  # noqa: E402
from eeyore_b200.datasets import XYDataset  # noqa: E402
from eeyore_b200.models.mlp import MLP, Hyperparameters  # noqa: E402
from eeyore_b200.samplers import AM, RAM, PowerPosteriorSampler  # noqa: E402

xor = XYDataset.from_eeyore("xor", dtype=torch.float64)
m = MLP(loss=loss_functions["binary_classification"], hparams=Hyperparameters([2, 3, 2, 1], 3 * [True], 3 * [torch.sigmoid]),
        dtype=torch.float64)
P = m.num_params()
m.prior = Normal(torch.zeros(P, dtype=torch.float64), 3 ** 0.5 * torch.ones(P, dtype=torch.float64))
loader = DataLoader(xor, batch_size=4)
g = torch.Generator().manual_seed(0)


def timed(make, iters, evals_per_iter, name):
    for _ in range(3):
        s = make()
        s.run(num_epochs=iters, num_burnin_epochs=iters)
    torch.cuda.synchronize()
    s = make()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s.run(num_epochs=iters, num_burnin_epochs=iters // 2)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{name}: {ms:.2f} ms for {iters} iterations -> {evals_per_iter * iters / ms * 1e3:.3e} log-target evaluations/s", flush=True)


C = 16384
th = torch.randn(C, P, dtype=torch.float64, generator=g) * 0.5
timed(lambda: RAM(m, theta0=th, dataloader=loader, seed=1), 200, C, f"RAM  {C} chains")
timed(lambda: AM(m, theta0=th, dataloader=loader, seed=1, t0=100, c=0.3, b=0.5), 200, C, f"AM   {C} chains (t0 = 100)")
E, K = 4096, 5
th = torch.randn(E, P, dtype=torch.float64, generator=g) * 0.5
timed(lambda: PowerPosteriorSampler(m, loader, [["MALA", {"step": 0.2}] for _ in range(K)], theta0=th, between_step=10, seed=2),
      200, E * K * (1 + 2 / 10), f"PowerPosterior {E} ensembles x {K} MALA levels")
