"""GPU parity: AM (eeyore/samplers/am.py) and RAM (eeyore/samplers/ram.py) -- SURVEY section 8f row 4 -- against the reference
goldens and the oracle, fed the reference's proposal noise and uniforms."""
import numpy as np
import pytest
import torch
from torch.distributions import Normal
from torch.utils.data import DataLoader

import oracle
from eeyore_b200.constants import loss_functions
from eeyore_b200.datasets import XYDataset
from eeyore_b200.models.mlp import MLP, Hyperparameters
from eeyore_b200.samplers import AM, RAM
from gpu_helpers import npy
from helpers import ARCHS, data_of, load, rel_err, spec_of

pytestmark = pytest.mark.gpu
S3 = 3 ** 0.5


def make(arch):
    a = ARCHS[arch]
    x, y = data_of(arch, np.float64)
    ds = XYDataset(torch.from_numpy(x), torch.from_numpy(y))
    nl = len(a["dims"]) - 1
    binary = a["data"] == "xor"
    acts = nl * [torch.sigmoid] if binary else (nl - 1) * [torch.sigmoid] + [None]
    loss = "binary_classification" if binary else "multiclass_classification"
    m = MLP(loss=loss_functions[loss], hparams=Hyperparameters(a["dims"], nl * [True], acts), dtype=torch.float64)
    P = m.num_params()
    m.prior = Normal(torch.zeros(P, dtype=torch.float64), S3 * torch.ones(P, dtype=torch.float64))
    return m, DataLoader(ds, batch_size=len(ds)), x, y, P


@pytest.mark.parametrize("name", ["am_xor221_f64", "am_xor2321_f64", "ram_xor221_f64", "ram_xor2321_f64"])
def test_reference_golden_single_chain(name):
    gd = load(name)
    arch = "221" if "221" in name else "2321"
    m, loader, _, _, P = make(arch)
    if name.startswith("am"):
        s = AM(m, theta0=torch.from_numpy(gd["theta0"]), dataloader=loader, l=float(gd["l"]), b=float(gd["b"]), c=float(gd["c"]),
               t0=int(gd["t0"]))
    else:
        s = RAM(m, theta0=torch.from_numpy(gd["theta0"]), dataloader=loader, a=float(gd["a"]), g=float(gd["g"]))
    s.set_noise_tape(torch.from_numpy(gd["z"]), torch.from_numpy(gd["u"]))
    s.run(num_epochs=int(gd["n_iters"]), num_burnin_epochs=int(gd["n_burnin"]))
    ch = s.get_chain()
    assert np.array_equal(np.array(ch.vals["accepted"], dtype=np.uint8), gd["accepted"])
    assert rel_err(npy(ch.get_samples()), gd["samples"]) < 1e-9
    assert np.allclose(npy(ch.get_target_vals()), gd["target_vals"], rtol=1e-9, atol=1e-11)
    factor = s.cov if name.startswith("am") else s.chol_cov
    assert rel_err(npy(factor), gd["final_factor"]) < 1e-7


@pytest.mark.parametrize("kind", ["am", "ram"])
@pytest.mark.parametrize("arch", ["2321", "433"])
def test_batched_chains_vs_oracle(kind, arch):
    """69 chains (ragged last block), continued over two run() calls."""
    m, loader, x, y, P = make(arch)
    spec = spec_of(arch)
    t0 = 3 * P                                        # enough distinct states for the plain covariance estimate
    C, T = 69, 3 * P + 30
    rng = np.random.default_rng(5)
    theta0 = rng.normal(size=(C, P)) * 0.5
    z = rng.normal(size=(T, C, P))
    cov0 = 0.05 * np.eye(P)
    if kind == "am":
        u = rng.uniform(size=(T, 2, C))
        ref = oracle.am_run(spec, x, y, np.zeros(P), np.full(P, S3), theta0, z, u, n_burnin=10, cov0=cov0, l=0.3, b=0.5, c=0.05, t0=t0)
        s = AM(m, theta0=torch.from_numpy(theta0), dataloader=loader, cov0=torch.from_numpy(cov0), l=0.3, b=0.5, c=0.05, t0=t0)
    else:
        u = rng.uniform(size=(T, C))
        ref = oracle.ram_run(spec, x, y, np.zeros(P), np.full(P, S3), theta0, z, u, n_burnin=10, cov0=cov0, a=0.3, g=0.65)
        s = RAM(m, theta0=torch.from_numpy(theta0), dataloader=loader, cov0=torch.from_numpy(cov0), a=0.3, g=0.65)
    s.set_noise_tape(torch.from_numpy(z), torch.from_numpy(u))
    s.run(num_epochs=40, num_burnin_epochs=10)
    s.run(num_epochs=T - 40, num_burnin_epochs=10)     # T - 40 more draws from iteration 40 on (counter, adaptive state, tape)
    ch = s.get_chain()
    assert np.array_equal(npy(ch.accepted_soa), ref["accepted"])
    got = np.transpose(npy(ch.get_samples()), (1, 0, 2))
    assert rel_err(got, ref["sample"]) < 1e-8
    key = "cov" if kind == "am" else "chol_cov"
    mine = npy(s.cov if kind == "am" else s.chol_cov)
    want = ref["final"][key] if kind == "am" else np.tril(ref["final"][key])
    assert rel_err(mine, want) < 1e-6
    assert 0.05 < ref["accepted"].mean() < 0.98


def test_philox_mode_and_errors():
    m, loader, x, y, P = make("221")
    th0 = torch.randn(256, P, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    a = RAM(m, theta0=th0, dataloader=loader, seed=3)
    a.run(num_epochs=300, num_burnin_epochs=100)
    b = RAM(m, theta0=th0, dataloader=loader, seed=3)
    b.run(num_epochs=300, num_burnin_epochs=100)
    sa = a.get_chain().get_samples()
    assert torch.equal(sa, b.get_chain().get_samples()) and torch.isfinite(sa).all()
    rate = a.acceptance_counts().double().mean().item() / 300
    assert 0.15 < rate < 0.35                          # RAM steers the acceptance rate to a = 0.234
    with pytest.raises(RuntimeError):                  # plain covariance estimate is singular right after t0 = 2
        s = AM(m, theta0=th0[:8], dataloader=loader, seed=1)
        s.run(num_epochs=30, num_burnin_epochs=0)
    with pytest.raises(ValueError):
        AM(m, theta0=th0[:8], dataloader=loader, transform=lambda c: c)
