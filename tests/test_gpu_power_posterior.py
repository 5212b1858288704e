"""GPU parity: PowerPosteriorSampler (eeyore/samplers/power_posterior_sampler.py) against the oracle and
the reference goldens, fed the reference's proposal noise, categorical neighbour draws and accept uniforms."""
import numpy as np
import pytest
import torch
from torch.distributions import Normal
from torch.utils.data import DataLoader

from eeyore_b200.constants import loss_functions
from eeyore_b200.datasets import XYDataset
from eeyore_b200.models.mlp import MLP, Hyperparameters
from eeyore_b200.samplers import PowerPosteriorSampler
from gpu_helpers import npy
from helpers import ARCHS, data_of, pp_setup, rel_err
from oracle.power_posterior import power_posterior_run

pytestmark = pytest.mark.gpu
S3 = 3 ** 0.5


def make(arch):
    a = ARCHS[arch]
    x, y = data_of(arch, np.float64)
    ds = XYDataset(torch.from_numpy(x), torch.from_numpy(y))
    nl = len(a["dims"]) - 1
    m = MLP(loss=loss_functions["binary_classification"], hparams=Hyperparameters(a["dims"], nl * [True], nl * [torch.sigmoid]),
            dtype=torch.float64)
    P = m.num_params()
    m.prior = Normal(torch.zeros(P, dtype=torch.float64), S3 * torch.ones(P, dtype=torch.float64))
    return m, DataLoader(ds, batch_size=len(ds)), x, y, P


@pytest.mark.parametrize("name", ["pp_221_mix", "pp_2321_mala"])
def test_reference_golden_single_ensemble(name):
    gd, spec, x, y, kinds, kwargs = pp_setup(name)
    arch = "221" if "221" in name else "2321"
    m, loader, _, _, P = make(arch)
    spec_samplers = [["MetropolisHastings", {}] if k == "mh" else ["MALA", {"step": kw["step"]}] for k, kw in zip(kinds, kwargs)]
    s = PowerPosteriorSampler(m, loader, spec_samplers, theta0=torch.from_numpy(gd["theta0"]), between_step=int(gd["between_step"]))
    assert np.allclose(s.temperature, gd["temperatures"])
    s.set_noise_tape(gd["z"], gd["u"], gd["j_tape"], gd["u_between"])
    s.run(num_epochs=int(gd["n_iters"]), num_burnin_epochs=int(gd["n_burnin"]))
    for k in range(len(kinds)):
        ch = s.get_chain(idx=k) if k else s.samplers[0].get_chain()
        assert rel_err(npy(ch.get_samples()), gd["samples"][k]) < 1e-10, k
        assert np.allclose(npy(ch.get_target_vals()), gd["target_vals"][k], rtol=1e-10, atol=1e-12)
    assert s.get_chain() is s.samplers[-1].get_chain()               # default indicator = the temperature-1 level
    assert rel_err(npy(s.samplers[1].current["sample"]), gd["final_sample"][1]) < 1e-10
    assert s.num_between_sweeps == gd["j_tape"].shape[0] and int(s.swap_count.sum()) > 0


def test_batched_ensembles_vs_oracle():
    """37 independent ensembles x 4 levels (MH and MALA mixed), sweeps every other iteration, random tapes."""
    m, loader, x, y, P = make("2321")
    from helpers import spec_of
    spec = spec_of("2321")
    E, K, T, bs, burn = 37, 4, 23, 2, 5
    rng = np.random.default_rng(7)
    theta0 = rng.normal(size=(E, P)) * 0.7
    z, u = rng.normal(size=(T, K, E, P)), rng.uniform(size=(T, K, E))
    nb = len([t for t in range(T) if t % bs == 0])
    jt = np.stack([np.stack([rng.choice([j for j in range(K) if j != i], size=E) for i in range(K)]) for _ in range(nb)])
    ub = rng.uniform(size=(nb, K, E))
    kinds, kwargs = ["mala", "mh", "mala", "mh"], [{"step": 0.3}, {}, {"step": 0.1}, {}]
    temps = [0.05, 0.2, 0.6, 1.0]
    ref = power_posterior_run(spec, x, y, np.zeros(P), np.full(P, S3), theta0, kinds, kwargs, z, u, jt, ub, temperatures=temps,
                              between_step=bs, b=0.5, n_burnin=burn)
    spec_samplers = [["MALA", {"step": 0.3}], ["MetropolisHastings", {}], ["MALA", {"step": 0.1}], ["MetropolisHastings", {}]]
    s = PowerPosteriorSampler(m, loader, spec_samplers, theta0=torch.from_numpy(theta0), temperature=temps, between_step=bs)
    s.set_noise_tape(z, u, jt, ub)
    s.run(num_epochs=T, num_burnin_epochs=burn)
    for k in range(K):
        ch = s.samplers[k].get_chain()
        got = npy(ch.get_samples())                                  # [E, n, P]
        assert rel_err(np.transpose(got, (1, 0, 2)), ref["sample"][k]) < 1e-10, k
        assert np.allclose(npy(ch.get_target_vals()).T, ref["target_val"][k], rtol=1e-10, atol=1e-12)
    assert np.array_equal(npy(s.swap_count), ref["swaps"].sum(axis=(0, 2)))
    assert 0 < ref["swaps"].sum() < ref["swaps"].size


def test_philox_mode_runs_and_validates():
    m, loader, x, y, P = make("221")
    th0 = torch.randn(64, P, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    spec_samplers = [["MALA", {"step": 0.2}] for _ in range(5)]
    a = PowerPosteriorSampler(m, loader, spec_samplers, theta0=th0, between_step=3, seed=11)
    a.run(num_epochs=40, num_burnin_epochs=10)
    b = PowerPosteriorSampler(m, loader, spec_samplers, theta0=th0, between_step=3, seed=11)
    b.run(num_epochs=40, num_burnin_epochs=10)
    sa, sb = a.get_chain().get_samples(), b.get_chain().get_samples()
    assert sa.shape == (64, 30, P) and torch.equal(sa, sb) and torch.isfinite(sa).all()
    assert a.num_between_sweeps == 14 and 0 < float(a.swap_rates().mean()) < 1
    with pytest.raises(ValueError):
        PowerPosteriorSampler(m, loader, [["HMC", {}], ["MALA", {}]], theta0=th0)
    with pytest.raises(ValueError):
        PowerPosteriorSampler(m, loader, spec_samplers, theta0=th0, temperature=[0.5, 1.0])


def test_file_storage_matches_list_storage(tmp_path):
    """PowerPosteriorSampler(storage='file') (power_posterior_sampler.py:57-63): chainN/<key>.csv per level, equal to the
    in-memory chains of the same run."""
    gd, spec, x, y, kinds, kwargs = pp_setup("pp_221_mix")
    m, loader, _, _, P = make("221")
    spec_samplers = [["MetropolisHastings", {}] if k == "mh" else ["MALA", {"step": kw["step"]}] for k, kw in zip(kinds, kwargs)]
    runs = {}
    for storage in ("list", "file"):
        s = PowerPosteriorSampler(m, loader, spec_samplers, theta0=torch.from_numpy(gd["theta0"]),
                                  between_step=int(gd["between_step"]), storage=storage, path=tmp_path)   # mode 'a': every launch appends
        s.set_noise_tape(gd["z"], gd["u"], gd["j_tape"], gd["u_between"])
        s.run(num_epochs=int(gd["n_iters"]), num_burnin_epochs=int(gd["n_burnin"]))
        runs[storage] = s
    for k in range(len(kinds)):
        mem = runs["list"].samplers[k].get_chain()
        disk = runs["file"].samplers[k].get_chain().to_chainlist(keys=["sample", "target_val"])
        assert (tmp_path / f"chain{k + 1}" / "sample.csv").exists()
        assert torch.equal(disk.get_samples(), mem.get_samples().cpu())
        assert torch.equal(disk.get_target_vals(), mem.get_target_vals().cpu())
